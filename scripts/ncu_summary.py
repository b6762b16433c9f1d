#!/usr/bin/env python
"""Summarise gpurun_out/prof_*.ncu-rep into profiles/<tag>_ncu_summary.json
(and print a markdown table).  Usage: python scripts/ncu_summary.py r01"""
import csv, glob, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "sm__inst_executed.avg.per_cycle_active": "ipc",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "ld_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "ld_requests",
    # L2 -> SM side (what bounds the gather kernels)
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors_from_sm",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "xbar_to_l1_read_bytes",
    "derived__lts__lts2xbar_bytes.sum": "l2_to_xbar_bytes",
    "derived__lts__lts2xbar_bytes.sum.per_second": "l2_to_xbar_bytes_per_s",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed": "l1_writeback_pct",
    "lts__t_sector_throughput_srcunit_tex.avg.pct_of_peak_sustained_elapsed": "l2_tex_sector_tput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.max": "sm_cycles",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "ns": 1e-3, "us": 1.0,
         "ms": 1e3, "Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6}
import re
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio$|"
                   r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio$")
out = {}
for rep in sorted(glob.glob(os.path.join(root, "gpurun_out", "prof_*.ncu-rep"))):
    name = os.path.basename(rep)[5:-8]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name}
    stalls = {}
    for i, h in enumerate(hdr):
        m = STALL.match(h)
        if m:
            try:
                stalls[m.group(1) or m.group(2)] = round(float(vals[i]), 3)
            except ValueError:
                pass
        if h in WANT:
            try:
                v = float(vals[i])
            except ValueError:
                continue
            v *= SCALE.get(unit[i], 1.0)
            d[WANT[h]] = v
    if stalls:
        # warps stalled per issued instruction, by reason (largest first)
        d["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    if "dram_read" in d and "dram_write" in d:
        d["dram_bytes"] = d["dram_read"] + d["dram_write"]
    if d.get("ld_requests"):
        d["sectors_per_request"] = d["ld_sectors"] / d["ld_requests"]
    out[name] = d
try:
    out["commit"] = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"],
                                   capture_output=True, text=True).stdout.strip()
except OSError:
    pass
os.makedirs(os.path.join(root, "profiles"), exist_ok=True)
path = os.path.join(root, "profiles", f"{tag}_ncu_summary.json")
json.dump(out, open(path, "w"), indent=1)
print("| kernel | dur us | DRAM MB (r+w) | L2 hit % | L1 hit % | occ % | regs | IPC | issue % | L2 tput % | DRAM tput % | sect/req |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, d in out.items():
    if not isinstance(d, dict):
        continue
    print(f"| {k} | {d.get('duration_us',0):.1f} | {d.get('dram_bytes',0)/1e6:.1f} | {d.get('l2_hit_pct',0):.1f} | "
          f"{d.get('l1_hit_pct',0):.1f} | {d.get('achieved_occupancy_pct',0):.1f} | {int(d.get('registers',0))} | "
          f"{d.get('ipc',0):.2f} | {d.get('issue_active_pct',0):.1f} | {d.get('l2_throughput_pct',0):.1f} | "
          f"{d.get('dram_throughput_pct',0):.1f} | {d.get('sectors_per_request',0):.1f} |")
print("wrote", path)
