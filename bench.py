#!/usr/bin/env python
"""bench.py -- headline benchmark of the embedding hot path on B200.

A "step" is one pass of the hot path over one synthetic batch:
    forward pool  ->  index transpose (row ids + sort + compressed remap)
                  ->  backward (compressed gradient)
on the manual_benchmark default workload of the reference (README.md:104):
10 M x 256 fp16 table, batch 65536, hotness 64, alpha 1.15, int32 indices,
fixed hotness, unweighted sum, compressed gradient, skip_grad_init.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  Stage times are CUDA-event times on the
launching stream with an L2 flush (write of a 512 MB buffer) before every
stage, so no stage starts with its inputs cached by the previous one --
the reference's protocol (benchmarks/manual_benchmark.cu:199-248).
Algorithmic bytes per stage are the reference's own formulas
(benchmarks/manual_benchmark.cu:250-261,340-354,444-473); see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------- workload
WORKLOADS = {
    # BASELINE.json configs[1]
    "C2": dict(num_categories=10_000_000, embed_width=256, batch_size=65536,
               hotness=64, alpha=1.15, dtype="f16", index="int32"),
    # BASELINE.json configs[0] (CPU reference check shape)
    "C1": dict(num_categories=1_048_576, embed_width=32, batch_size=1024,
               hotness=8, alpha=0.0, dtype="f32", index="int32"),
    # BASELINE.json configs[4]: ONE table sharded by rows over the ranks (strong
    # scaling: the global problem is fixed; N > 1 only)
    "C5": dict(num_categories=400_000_000, embed_width=128, batch_size=262144,
               hotness=64, alpha=1.15, dtype="f16", index="int32", global_problem=True),
    # large-batch shape: same nnz as C2, but grad_y (524 288 bags = 268 MB) does
    # not fit in L2 (the shape of one rank's local backward in the 8-way
    # row-sharded C2)
    "bigbatch": dict(num_categories=10_000_000, embed_width=256, batch_size=524288,
                     hotness=8, alpha=1.15, dtype="f16", index="int32"),
    # small shape for smoke runs
    "tiny": dict(num_categories=100_000, embed_width=256, batch_size=4096,
                 hotness=64, alpha=1.15, dtype="f16", index="int32"),
}


def workload_string(name, cfg):
    """The same text in the line of this library and in the reference arm's."""
    return (f"manual_benchmark default ({name}): {cfg['num_categories']}x{cfg['embed_width']} "
            f"{cfg['dtype']}, batch {cfg['batch_size']}, hotness {cfg['hotness']}, "
            f"alpha {cfg['alpha']}, {cfg['index']} indices, sum, compressed grad")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def stage_bytes(cfg, nnz, num_unique):
    """Algorithmic bytes per stage, the reference's accounting."""
    es = 2 if cfg["dtype"] in ("f16", "bf16") else 4
    isz = 4 if cfg["index"] == "int32" else 8
    w, b = cfg["embed_width"], cfg["batch_size"]
    fwd = es * w * (nnz + b)                                   # manual_benchmark.cu:256-260
    tr = nnz * isz * (1 + 3)                                   # :340-354, compressed
    bwd_dram = es * w * num_unique + 2 * isz * nnz + es * w * b  # :451-467
    bwd_l2 = bwd_dram + es * w * nnz                           # :468-473
    return dict(forward=fwd, transpose=tr, backward=bwd_l2, backward_dram=bwd_dram)


# ------------------------------------------------------------ CPU baselines
def cpu_reference_pass(cfg, wl_indices, table_host, grad_y_host, sample_bags,
                       threads, kind):
    """One pass of the reference CPU path on the first `sample_bags` bags.
    Returns (seconds per stage dict, sample nnz)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cpu_lib
    lib = cpu_lib.CpuLib(kind)
    hot, w = cfg["hotness"], cfg["embed_width"]
    nnz = sample_bags * hot
    idx = np.ascontiguousarray(wl_indices[:nnz])
    t = {}
    ret = cpu_lib.empty_like_dt((sample_bags, w), cpu_lib.dt_code(table_host))
    bounds = np.linspace(0, sample_bags, threads + 1).astype(int)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda i: lib.forward(
            table_host, idx, None, None, sample_bags, hot, cpu_lib.SUM,
            embed_width=w, ret=ret, sample_begin=int(bounds[i]),
            sample_end=int(bounds[i + 1])), range(threads)))
    t["forward"] = time.perf_counter() - t0
    # transpose: single-threaded in the reference (std::sort of tuples)
    t0 = time.perf_counter()
    rows = lib.extract_row_ids_fixed(sample_bags, hot, idx.dtype)
    t_idx, t_sid, _ = lib.transpose(rows, idx, None)
    remapped = lib.compressed_grad_indices(t_idx)
    t["transpose"] = time.perf_counter() - t0
    num_unique = int(remapped[-1]) + 1
    grad = cpu_lib.empty_like_dt((num_unique, w), cpu_lib.dt_code(grad_y_host))
    inv = np.zeros(num_unique, idx.dtype)
    # backward sliced at run boundaries across threads (pointer offsets only)
    cuts = [0]
    for i in range(1, threads):
        c = int(nnz * i / threads)
        while c < nnz and c > 0 and t_idx[c] == t_idx[c - 1]:
            c += 1
        cuts.append(max(c, cuts[-1]))
    cuts.append(nnz)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda i: lib.backward(
            grad_y_host, w, num_unique, t_idx, t_sid, remapped, None,
            skip_grad_init=True, grad_embedding=grad, inverse_mapping=inv,
            nz_begin=cuts[i], nz_end=cuts[i + 1]) if cuts[i + 1] > cuts[i] else None,
            range(threads)))
    t["backward"] = time.perf_counter() - t0
    return t, nnz


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_kind():
    from oracle import cpu_lib
    if cpu_lib.ref_available():
        return "ref", "reference"
    cpu_lib.build(ref=False)
    return "oracle", "port"


def make_host_inputs(cfg, sample_bags, seed=1234):
    """Host arrays for the CPU legs: full index list + table + grad_y sample."""
    from cuembed_b200 import datagen
    idt = np.int32 if cfg["index"] == "int32" else np.int64
    # The index list is deterministic (fixed seed); cache it between runs of the
    # same box session to keep parameter sweeps short.
    key = "_".join(str(cfg[k]) for k in ("num_categories", "batch_size", "hotness", "alpha", "index"))
    cache = os.path.join("/tmp", f"cuembed_b200_wl_{key}_{seed}.npy")
    if os.path.exists(cache):
        idx = np.load(cache)
        return datagen.Workload(cfg["num_categories"], cfg["embed_width"], cfg["batch_size"],
                                cfg["hotness"], idx, None, None, int(idx.shape[0]))
    wl = datagen.make_workload(cfg["num_categories"], cfg["embed_width"],
                               cfg["batch_size"], cfg["hotness"],
                               alpha=cfg["alpha"], seed=seed, index_dtype=idt)
    try:
        np.save(cache, wl.indices)
    except OSError:
        pass
    return wl



# ------------------------------------------------- measured gather ceilings
def measure_gather_ceilings(torch, dev, table, grad_y, indices, t_sid, flush, reps=5):
    """Roofline denominators measured in this run (csrc/microbench.cu): the
    product kernels' access shape with no arithmetic.
      l2_random   : uniform random rows of the L2-resident grad_y buffer
                    (33.6 MB at C2) -> the L2 -> SM gather ceiling
      dram_random : uniform random rows of the whole table -> DRAM gather ceiling
      fwd_stream / bwd_stream : the step's own index streams (power-law lookups
                    over the table; sorted sample ids over grad_y), L2 flushed
                    first like the stages themselves
    All in GB/s of gathered row bytes."""
    import ctypes
    from cuembed_b200 import _lib
    lib = _lib.load()
    row_bytes = table.shape[1] * table.element_size()
    if row_bytes not in (128, 256, 512):
        return None
    sink = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def run(buf, rows, flush_first):
        n = rows.numel()
        best = None
        for it in range(reps + 1):
            if flush_first:
                flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = lib.cuembed_microbench_gather(buf.data_ptr(), row_bytes, rows.data_ptr(), n, 0,
                                               sink.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            if rc != 0:
                return None
            ms = e0.elapsed_time(e1)
            if it > 0:
                best = ms if best is None else min(best, ms)
        return {"ms": round(best, 4), "GBps": round(n * row_bytes / (best * 1e-3) / 1e9, 1)}

    g = torch.Generator(device=dev)
    g.manual_seed(42)
    n = indices.numel()
    rnd_small = torch.randint(0, grad_y.shape[0], (n,), generator=g, device=dev, dtype=torch.int32)
    rnd_big = torch.randint(0, table.shape[0], (n,), generator=g, device=dev, dtype=torch.int32)
    return {
        "l2_random": run(grad_y, rnd_small, False),
        "dram_random": run(table, rnd_big, True),
        "fwd_stream": run(table, indices.to(torch.int32), True),
        "bwd_stream": run(grad_y, t_sid.to(torch.int32), True),
        "row_bytes": row_bytes, "rows_gathered": n,
        "how": "cuembed_microbench_gather: lane group per row, 16-byte loads, 8 rows in flight, "
               "persistent grid, XOR only; best of %d, CUDA events" % reps,
    }


def reference_gpu_same_box(args, torch, dev, tdt, cfg, table, indices, grad_y, flush, ours_ms):
    """The bar of SURVEY.md section 0: the reference's OWN GPU kernels (its headers
    compiled unchanged for sm_100 into oracle/_ref/libcuembed_refgpu.so by
    oracle/Makefile `refgpu`), timed here on the same device-resident inputs
    with the same protocol as `value` (L2 flush before every stage, CUDA
    events, each stage replayed from a CUDA graph unless --no-graphs), plus a
    cross-check of the results.  Bench infrastructure; outside `value`."""
    import ctypes
    path = os.path.join(ROOT, "oracle", "_ref", "libcuembed_refgpu.so")
    if not os.path.exists(path) or cfg["dtype"] not in ("f16", "f32") or cfg["index"] != "int32":
        return {"unavailable": "oracle/_ref/libcuembed_refgpu.so not built or dtype unsupported"}
    ref = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    dtc = {"f16": 1, "f32": 0}[cfg["dtype"]]
    w, batch, hot = cfg["embed_width"], cfg["batch_size"], cfg["hotness"]
    nnz = batch * hot
    P = lambda t: vp(t.data_ptr()) if t is not None else None  # noqa: E731
    S = lambda: vp(torch.cuda.current_stream().cuda_stream)    # noqa: E731
    for fn in ("forward", "extract_row_ids_fixed", "transpose", "compressed_grad_indices", "backward"):
        getattr(ref, "refgpu_" + fn).restype = ctypes.c_int
    out = torch.empty(batch, w, dtype=tdt, device=dev)
    row_ids = torch.empty(nnz, dtype=torch.int32, device=dev)
    t_idx = torch.empty(nnz, dtype=torch.int32, device=dev)
    t_sid = torch.empty(nnz, dtype=torch.int32, device=dev)
    rem = torch.empty(nnz, dtype=torch.int32, device=dev)
    lw, lw2 = ctypes.c_size_t(0), ctypes.c_size_t(0)
    ref.refgpu_transpose(None, None, None, dtc, nnz, 0, None, None, None, None, ctypes.byref(lw), S())
    ref.refgpu_compressed_grad_indices(None, 0, nnz, None, None, ctypes.byref(lw2), S())
    lwork = max(lw.value, lw2.value)
    work = torch.empty(lwork, dtype=torch.uint8, device=dev)

    def fwd():
        ref.refgpu_forward(P(table), dtc, w, P(indices), 0, None, 0, None, batch, hot, 0, 0,
                           P(out), dtc, S())

    def tr():
        ref.refgpu_extract_row_ids_fixed(batch, hot, P(row_ids), 0, S())
        l1 = ctypes.c_size_t(lwork)
        ref.refgpu_transpose(P(row_ids), P(indices), None, dtc, nnz, 0, P(t_idx), P(t_sid), None,
                             P(work), ctypes.byref(l1), S())
        l2 = ctypes.c_size_t(lwork)
        ref.refgpu_compressed_grad_indices(P(t_idx), 0, nnz, P(rem), P(work), ctypes.byref(l2), S())

    fwd()
    tr()
    torch.cuda.synchronize()
    nu = int(rem[-1].item()) + 1
    grad = torch.zeros(nu, w, dtype=tdt, device=dev)
    inv = torch.empty(nu, dtype=torch.int32, device=dev)

    def bwd():
        ref.refgpu_backward(P(grad_y), dtc, w, nu, nnz, 0, P(t_idx), P(t_sid), P(rem), None, 1,
                            P(grad), P(inv), S())

    res = {}
    mode = "direct launches"
    for name, fn in (("forward", fwd), ("transpose", tr), ("backward", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        run = fn
        if not args.no_graphs:
            try:
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    fn()
                run = g_.replay
                run()
                torch.cuda.synchronize()
                mode = "one CUDA graph per stage (captured once, replayed)"
            except Exception:  # noqa: BLE001 -- CUB may allocate: fall back to direct launches
                torch.cuda.synchronize()
                run = fn
        tot = 0.0
        for _ in range(args.steps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[name] = tot / args.steps
    grad.zero_()
    bwd()
    torch.cuda.synchronize()
    total = sum(res.values())
    return {
        "what": "the reference's own kernels (headers compiled unchanged for sm_100, CUB sort) "
                "on the same B200, same inputs, same protocol",
        "launch": mode,
        "reference_gpu_ms": {k: round(v, 4) for k, v in res.items()},
        "ours_ms": {k: round(v, 4) for k, v in ours_ms.items()},
        "speedup": {k: round(res[k] / ours_ms[k], 3) for k in res},
        "total_ms": {"reference_gpu": round(total, 4), "ours": round(sum(ours_ms.values()), 4)},
        "speedup_total": round(total / sum(ours_ms.values()), 3),
        "outputs": {"out": out, "t_idx": t_idx, "t_sid": t_sid, "rem": rem, "grad": grad, "inv": inv},
    }


# ------------------------------------------------------------------ GPU arm
def run_e2e(args, ce, torch, dev, tdt, idt, cfg, idx_host, table, num_unique, work, bwork,
            fused_sgd=False):
    """End to end through the public API with HOST buffers: every step copies
    its indices and grad_y in from pinned memory and reads its pooled output,
    compressed gradient and row list back to pinned memory.

    The steps are software-pipelined the way a host-side caller of an
    asynchronous GPU library would: three streams (copy-in, kernels, copy-out)
    and two sets of device buffers, so the D2H of step i overlaps the H2D and
    the kernels of step i+1 (PCIe is full duplex).  Every step still moves all
    of its own bytes and the caller still reads num_unique back before it
    launches the backward (the reference's protocol,
    benchmarks/manual_benchmark.cu:392-394)."""
    w, batch, hot, rows = (cfg["embed_width"], cfg["batch_size"], cfg["hotness"],
                           cfg["num_categories"])
    nnz = batch * hot
    table_es = 2 if cfg["dtype"] in ("f16", "bf16") else 4
    g = torch.Generator(device=dev)
    g.manual_seed(654321)
    gy_host = torch.randint(-10, 11, (batch, w), generator=g, device=dev).to(tdt).cpu().pin_memory()
    s_in, s_k, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))

    class Slot:
        def __init__(self):
            self.indices = torch.empty(nnz, dtype=idt, device=dev)
            self.grad_y = torch.empty(batch, w, dtype=tdt, device=dev)
            self.out = torch.empty(batch, w, dtype=tdt, device=dev)
            self.row_ids = torch.empty(nnz, dtype=idt, device=dev)
            self.t_idx = torch.empty(nnz, dtype=idt, device=dev)
            self.t_sid = torch.empty(nnz, dtype=idt, device=dev)
            self.remapped = torch.empty(nnz, dtype=idt, device=dev)
            self.grad = torch.empty(num_unique, w, dtype=tdt, device=dev)
            self.inv = torch.empty(num_unique, dtype=idt, device=dev)
            self.out_host = torch.empty(batch, w, dtype=tdt).pin_memory()
            self.grad_host = torch.empty(num_unique, w, dtype=tdt).pin_memory()
            self.inv_host = torch.empty(num_unique, dtype=idt).pin_memory()
            self.nu_host = torch.empty(1, dtype=idt).pin_memory()
            self.in_done = torch.cuda.Event()
            self.k_done = torch.cuda.Event()
            self.out_done = torch.cuda.Event()

    slots = [Slot(), Slot()]

    def e2e_step(i):
        s = slots[i & 1]
        with torch.cuda.stream(s_in):
            s_in.wait_event(s.k_done)      # kernels of step i-2 are done with the inputs
            s.indices.copy_(idx_host, non_blocking=True)
            s.grad_y.copy_(gy_host, non_blocking=True)
            s.in_done.record(s_in)
        with torch.cuda.stream(s_k):
            s_k.wait_event(s.in_done)
            s_k.wait_event(s.out_done)     # step i-2's results have left the device
            ce.EmbeddingForward(table, w, s.indices, None, None, batch, hot,
                                ce.CombineMode.kSum, s.out)
            ce.TransposeFixed(s.indices, None, batch, hot, s.t_idx, s.t_sid, None, work)
            nu = 0
            if fused_sgd:
                # the gradient is consumed on the device: SGD step on the table
                # rows (lr = 0 keeps the benchmark table's bits)
                ce.EmbeddingBackwardUpdate(s.grad_y, w, nnz, s.t_idx, s.t_sid, None,
                                           ce.OPT_SGD, 0.0, table, work=bwork)
            else:
                ce.ComputeCompressedGradIndices(s.t_idx, nnz, s.remapped, work)
                s.nu_host.copy_(s.remapped[-1:], non_blocking=True)
                s_k.synchronize()          # the caller sizes the gradient
                nu = int(s.nu_host.item()) + 1
                ce.EmbeddingBackward(s.grad_y, w, nu, nnz, s.t_idx, s.t_sid, s.remapped,
                                     None, True, s.grad, s.inv, work=bwork)
            s.k_done.record(s_k)
        with torch.cuda.stream(s_out):
            s_out.wait_event(s.k_done)
            s.out_host.copy_(s.out, non_blocking=True)
            if not fused_sgd:
                s.grad_host[:nu].copy_(s.grad[:nu], non_blocking=True)
                s.inv_host[:nu].copy_(s.inv[:nu], non_blocking=True)
            s.out_done.record(s_out)

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s_in)
    for i in range(args.steps):
        e2e_step(i)
    s_out.synchronize()                    # the last step's results are on the host
    e1.record(s_out)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = max(e0.elapsed_time(e1) / args.steps, wall_ms)
    isz = idx_host.element_size()
    h2d = nnz * isz + batch * w * table_es
    d2h = batch * w * table_es + num_unique * w * table_es + num_unique * isz + isz
    if fused_sgd:
        d2h = batch * w * table_es
    return e2e_ms, h2d, d2h


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import cuembed_b200 as ce

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # With NCCL_DEBUG set NCCL prints its version banner on STDOUT, in front of
        # the one JSON line this program owes its caller.  File descriptor 1 is
        # pointed at stderr for everything that is not ours; the JSON line goes to
        # the original stdout (sys.stdout is rebound to it at the end of the run).
        sys.stdout.flush()
        real_stdout = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = real_stdout
        dist.init_process_group("nccl", device_id=dev)
        from benchmarks import sharded_bench
        return sharded_bench.run(args, rank, local_rank, world)

    cfg = WORKLOADS[args.workload]
    tdt = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}[cfg["dtype"]]
    idt = torch.int32 if cfg["index"] == "int32" else torch.int64
    rows, w, batch, hot = (cfg["num_categories"], cfg["embed_width"],
                           cfg["batch_size"], cfg["hotness"])
    nnz = batch * hot

    wl = make_host_inputs(cfg, batch)
    # table U(-1,1), generated on the device (5 GB for C2), seeded
    g = torch.Generator(device=dev)
    g.manual_seed(123456)
    table = torch.empty(rows, w, dtype=tdt, device=dev)
    chunk = 1 << 20
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        table[r0:r1] = (torch.rand(r1 - r0, w, generator=g, device=dev) * 2 - 1).to(tdt)
    g.manual_seed(654321)
    grad_y = torch.randint(-10, 11, (batch, w), generator=g, device=dev).to(tdt)

    idx_host = torch.from_numpy(wl.indices).pin_memory()
    indices = idx_host.to(dev, non_blocking=True)
    out = torch.empty(batch, w, dtype=tdt, device=dev)
    row_ids = torch.empty(nnz, dtype=idt, device=dev)
    t_idx = torch.empty(nnz, dtype=idt, device=dev)
    t_sid = torch.empty(nnz, dtype=idt, device=dev)
    remapped = torch.empty(nnz, dtype=idt, device=dev)
    lwork = max(ce.Transpose(row_ids, indices, None, nnz, None, None, None, None),
                ce.ComputeCompressedGradIndices(indices, nnz, None, None))
    work = torch.empty(lwork, dtype=torch.uint8, device=dev)
    bwork = torch.empty(ce.backward_workspace_bytes(tdt, w, nnz, idt),
                        dtype=torch.uint8, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def forward():
        ce.EmbeddingForward(table, w, indices, None, None, batch, hot,
                            ce.CombineMode.kSum, out)

    def transpose():
        # fixed-hotness COO: the sample ids (position / hotness) are synthesised in
        # the first sort pass (cuembed_transpose_fixed) instead of being written
        # by ExtractRowIdsFromFixed and read back; --separate-row-ids times the
        # reference's three-call sequence
        if args.separate_row_ids:
            ce.ExtractRowIdsFromFixed(batch, hot, row_ids)
            ce.Transpose(row_ids, indices, None, nnz, t_idx, t_sid, None, work)
        else:
            ce.TransposeFixed(indices, None, batch, hot, t_idx, t_sid, None, work)
        ce.ComputeCompressedGradIndices(t_idx, nnz, remapped, work)

    # first pass to learn num_unique (the caller reads remapped.back()+1 on the
    # host, benchmarks/manual_benchmark.cu:392-394)
    forward()
    transpose()
    num_unique = int(remapped[-1].item()) + 1
    grad = torch.zeros(num_unique, w, dtype=tdt, device=dev)
    inv = torch.empty(num_unique, dtype=idt, device=dev)

    def backward():
        ce.EmbeddingBackward(grad_y, w, num_unique, nnz, t_idx, t_sid, remapped,
                             None, True, grad, inv, work=bwork)

    stages = [("forward", forward), ("transpose", transpose), ("backward", backward)]
    stream = torch.cuda.current_stream()

    # Each stage is a fixed sequence of launches on fixed buffers (transpose: 9
    # kernels, some of them 3-6 us long), so it is captured once into a CUDA
    # graph and replayed: the kernels then follow each other without waiting
    # for the host to enqueue the next one.  The library never synchronises or
    # allocates on this path (explicit workspaces), which is what makes it
    # capturable.  --no-graphs launches directly.
    launch_mode = "direct launches"
    graph_kernels = 0
    if not args.no_graphs:
        try:
            for _ in range(2):  # warm the per-kernel attribute caches first
                for _, fn in stages:
                    fn()
            torch.cuda.synchronize()
            graphs = []
            graph_kernels = 0  # kernels recorded into the graphs = per step
            for name, fn in stages:
                g_ = torch.cuda.CUDAGraph()
                c0_ = ce.launch_count()
                with torch.cuda.graph(g_):
                    fn()
                graph_kernels += ce.launch_count() - c0_
                graphs.append((name, g_.replay))
            torch.cuda.synchronize()
            stages = graphs
            stream = torch.cuda.current_stream()
            launch_mode = "one CUDA graph per stage (captured once, replayed)"
        except Exception as e:  # noqa: BLE001 -- report and fall back
            launch_mode = f"direct launches (graph capture failed: {type(e).__name__})"
            torch.cuda.synchronize()

    def one_step(times=None):
        for name, fn in stages:
            flush.fill_(1)  # L2 flush: 512 MB write > 126 MB L2
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            if times is not None:
                times.append((name, e0, e1))

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ce.launch_count()
    events = []
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(events)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = ce.launch_count() - launches0
    if launch_mode.startswith("one CUDA graph"):
        # replays do not pass through the library's host code: every replayed
        # step runs the kernels that were recorded at capture time
        launches = graph_kernels * args.steps
    per_stage = {"forward": 0.0, "transpose": 0.0, "backward": 0.0}
    for name, e0, e1 in events:
        per_stage[name] += e0.elapsed_time(e1)
    for k in per_stage:
        per_stage[k] /= args.steps
    ms_per_step = sum(per_stage.values())

    # ---- extra (not part of `value`): the fused alternative to transpose +
    # backward, SURVEY.md 8(f) f3: sort without the compressed-index pass, then
    # backward fused with an SGD step on the table rows (lr = 0 keeps the table
    # bits, every instruction still runs).
    extras = {}
    if not args.no_extras:
        def transpose_nc():
            ce.TransposeFixed(indices, None, batch, hot, t_idx, t_sid, None, work)

        def update():
            ce.EmbeddingBackwardUpdate(grad_y, w, nnz, t_idx, t_sid, None, ce.OPT_SGD,
                                       0.0, table, work=bwork)

        ex_t = {"transpose_no_compress": 0.0, "backward_sgd_update": 0.0}
        ex_ev = []
        for it in range(3 + args.steps):
            for name, fn in (("transpose_no_compress", transpose_nc),
                             ("backward_sgd_update", update)):
                flush.fill_(1)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                if it >= 3:
                    ex_ev.append((name, e0, e1))
        torch.cuda.synchronize()
        for name, e0, e1 in ex_ev:
            ex_t[name] += e0.elapsed_time(e1) / args.steps
        extras["fused_sgd_step"] = {
            "ms": {k: round(v, 4) for k, v in ex_t.items()},
            "ms_total_with_forward": round(per_stage["forward"] + sum(ex_t.values()), 4),
            "note": "forward + sort + backward fused with the SGD step on the table "
                    "(no compressed indices, no gradient round trip); compare with "
                    "ms_per_step, which leaves the gradient for a separate optimizer"}
        transpose()  # restore remapped/t_idx for the e2e leg

    # ---- extra: the reference's own GPU kernels on this box (the bar to beat)
    # and the measured gather ceilings (the roofline denominators)
    ceilings = None
    if not args.no_extras:
        rg = reference_gpu_same_box(args, torch, dev, tdt, cfg, table, indices, grad_y, flush,
                                    per_stage)
        if "outputs" in rg:
            o = rg.pop("outputs")
            grad.zero_()
            backward()
            torch.cuda.synchronize()
            same = (grad == o["grad"])
            rg["cross_check"] = {
                "forward_equal": bool(torch.equal(out, o["out"])),
                "transpose_indices_equal": bool(torch.equal(t_idx, o["t_idx"])),
                "transpose_sample_ids_equal": bool(torch.equal(t_sid, o["t_sid"])),
                "remapped_equal": bool(torch.equal(remapped, o["rem"])),
                "inverse_mapping_equal": bool(torch.equal(inv, o["inv"])),
                "backward_frac_equal": round(float(same.float().mean().item()), 7),
                "backward_max_abs_diff": float((grad.float() - o["grad"].float()).abs().max().item()),
                "note": "the reference backward accumulates in the gradient type (fp16) with "
                        "atomics at block edges; this library accumulates in fp32 in a fixed order",
            }
            del o
        extras["reference_gpu_same_box"] = rg
        ceilings = measure_gather_ceilings(torch, dev, table, grad_y, indices, t_sid, flush)
        extras["gather_ceilings"] = ceilings
        # ---- the hot-row policy of north_star, measured: forward with a
        # shared-memory cache of the rows hit >= nnz / 4096 times in this batch
        # (list built on the device from the transposed indices)
        cap = ce.forward_hot_capacity(tdt, w)
        if cap > 0 and idt == torch.int32:
            hot_rows, hot_count = ce.HotRowsFromSorted(t_idx, nnz, max(2, nnz // 4096), cap)
            out_hot = torch.empty_like(out)

            def forward_hot():
                ce.EmbeddingForwardHot(table, w, indices, None, None, batch, hot,
                                       ce.CombineMode.kSum, out_hot, hot_rows, hot_count)

            def hot_list():
                ce.HotRowsFromSorted(t_idx, nnz, max(2, nnz // 4096), cap)

            tms = {}
            for name, fn in (("forward_hot", forward_hot), ("hot_list", hot_list)):
                for _ in range(3):
                    fn()
                tot = 0.0
                for _ in range(args.steps):
                    flush.fill_(1)
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    fn()
                    e1.record(stream)
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
                tms[name] = tot / args.steps
            forward()
            torch.cuda.synchronize()
            extras["forward_hot_row_cache"] = {
                "ms": round(tms["forward_hot"], 4), "plain_forward_ms": round(per_stage["forward"], 4),
                "hot_list_ms": round(tms["hot_list"], 4),
                "rows_cached": int(min(int(hot_count.item()), cap)), "capacity": cap,
                "min_count": max(2, nnz // 4096),
                "bit_identical_to_plain_forward": bool(torch.equal(out_hot, out)),
                "note": "cuembed_forward_hot: rows hit >= min_count times kept in shared memory "
                        "(one 1024-thread CTA per SM, hash probe per index); list from the "
                        "transposed indices of the same batch"}

    e2e_ms, h2d, d2h = float("nan"), 0, 0
    if not args.no_e2e:
        e2e_ms, h2d, d2h = run_e2e(args, ce, torch, dev, tdt, idt, cfg, idx_host, table,
                                   num_unique, work, bwork)
        if not args.no_extras:
            f_ms, f_h2d, f_d2h = run_e2e(args, ce, torch, dev, tdt, idt, cfg, idx_host, table,
                                         num_unique, work, bwork, fused_sgd=True)
            extras["e2e_fused_sgd_step"] = {
                "value": round(nnz / (f_ms * 1e-3), 1), "unit": "lookups/s",
                "ms_per_step": round(f_ms, 4), "h2d_bytes_per_step": int(f_h2d),
                "d2h_bytes_per_step": int(f_d2h),
                "note": "same host-buffer protocol, but the step ends in the fused SGD "
                        "update of the table in HBM, so only the pooled output returns "
                        "to the host (the headline e2e also ships the 293 MB compressed "
                        "gradient and is PCIe-bound)"}
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel.  Both gather kernels are bound by
    # the rate at which L2 delivers scattered rows to the SMs (power-law
    # indices: most row reads are L2 hits, the backward's all are), so the
    # denominator is the L2 -> SM gather ceiling MEASURED IN THIS RUN
    # (extras.gather_ceilings.l2_random); `dram_frac` puts the DRAM-level bytes
    # of the same kernel against the measured HBM copy bandwidth.
    hbm_peak, hbm_src = measured_peaks()
    by = stage_bytes(cfg, nnz, num_unique)
    dominant = max(("forward", "backward"), key=lambda k: per_stage[k])
    kernel_name = {"forward": "FwdPoolKernel", "backward": "BwdWarpKernel"}[dominant]
    achieved = by[dominant] / (per_stage[dominant] * 1e-3) / 1e9
    es_ = 2 if cfg["dtype"] in ("f16", "bf16") else 4
    isz_ = 4 if cfg["index"] == "int32" else 8
    dram_bytes = {"forward": es_ * w * (num_unique + batch) + isz_ * nnz,   # compulsory
                  "backward": by["backward_dram"]}
    if ceilings is not None and ceilings.get("l2_random"):
        peak, bound = float(ceilings["l2_random"]["GBps"]), "l2"
        peak_src = ("measured in this run: L2-resident random-row gather, "
                    "cuembed_microbench_gather (extras.gather_ceilings.l2_random)")
    else:
        peak, bound, peak_src = hbm_peak, "hbm", hbm_src
    # DRAM traffic of that kernel: `ncu --set full` capture of this workload,
    # per launch (a separate run; profiles/*_ncu_summary.json says which commit)
    traffic, traffic_src = None, None
    pdir = os.path.join(ROOT, "profiles")
    if args.workload == "C2" and os.path.isdir(pdir):
        summaries = sorted(f for f in os.listdir(pdir) if f.endswith("_ncu_summary.json"))
        if summaries:
            with open(os.path.join(pdir, summaries[-1])) as f:
                summ = json.load(f)
            if kernel_name in summ and "dram_bytes" in summ[kernel_name]:
                traffic = int(summ[kernel_name]["dram_bytes"])
                traffic_src = (f"profiles/{summaries[-1]} (ncu --set full, separate run, "
                               f"commit {summ.get('commit', 'unrecorded')})")
    roofline = {"bound": bound, "kernel": kernel_name, "stage": dominant,
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": by[dominant],
                "dram_level_bytes_per_launch": dram_bytes[dominant],
                "dram_frac": round(dram_bytes[dominant] / (per_stage[dominant] * 1e-3) / 1e9
                                   / hbm_peak, 4),
                "hbm_peak": hbm_peak, "hbm_peak_source": hbm_src}
    peak = hbm_peak
    agg_bytes = by["forward"] + by["transpose"] + by["backward"]
    stage_report = {k: {"ms": round(per_stage[k], 4),
                        "lookups_per_s": round(nnz / (per_stage[k] * 1e-3), 1),
                        "algo_GBps": round(by[k] / (per_stage[k] * 1e-3) / 1e9, 1),
                        "frac_of_hbm_peak": round(by[k] / (per_stage[k] * 1e-3) / 1e9 / peak, 4)}
                    for k in per_stage}
    if ceilings is not None:
        # the stage's own index stream pulled through the same access shape with
        # no arithmetic and no stores: what the memory system allows this schedule
        for st_, key_ in (("forward", "fwd_stream"), ("backward", "bwd_stream")):
            if ceilings.get(key_):
                stage_report[st_]["gather_only_ms"] = ceilings[key_]["ms"]
                stage_report[st_]["frac_of_gather_only"] = round(
                    ceilings[key_]["ms"] / per_stage[st_], 4)
    stage_report["aggregate"] = {
        "algo_GBps": round(agg_bytes / (ms_per_step * 1e-3) / 1e9, 1),
        "frac_of_hbm_peak": round(agg_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4)}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if not args.no_cpu_baseline:
        kind_lib, kind = cpu_kind()
        sample_bags = min(batch, args.cpu_sample_bags)
        threads = host_threads()
        table_host = table.cpu().numpy() if tdt != torch.bfloat16 else None
        gyh = grad_y.cpu().numpy() if tdt != torch.bfloat16 else None
        if table_host is None:
            from oracle.cpu_lib import Bf16
            table_host = Bf16(table.view(torch.int16).cpu().numpy().view(np.uint16))
            gyh = Bf16(grad_y.view(torch.int16).cpu().numpy().view(np.uint16))
        t, s_nnz = cpu_reference_pass(cfg, wl.indices, table_host, gyh,
                                      sample_bags, threads, kind_lib)
        total = sum(t.values())
        cpu = {"value": round(s_nnz / total, 1), "unit": "lookups/s",
               "cores": threads, "kind": kind,
               "sample": f"first {sample_bags} of {batch} bags ({s_nnz} lookups), "
                         f"fwd+transpose+bwd once; fwd and bwd sliced over {threads} "
                         f"threads, transpose single-threaded as in the reference",
               "stage_s": {k: round(v, 4) for k, v in t.items()}}

    line = {
        "metric": "lookups/s (fwd + transpose + bwd, compressed grad)",
        "value": round(nnz / (ms_per_step * 1e-3), 1),
        "unit": "lookups/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": workload_string(args.workload, cfg),
                   "l2": "flushed before every stage (512 MB write)",
                   "launch": launch_mode,
                   "transpose": ("ExtractRowIdsFromFixed + Transpose + ComputeCompressedGradIndices"
                                 if args.separate_row_ids else
                                 "cuembed_transpose_fixed (row ids synthesised in the first sort "
                                 "pass) + ComputeCompressedGradIndices"),
                   "frac_of_hbm_peak": "SURVEY 8(d) accounting (the reference's algorithmic bytes / "
                                       "stage time / measured HBM copy bandwidth): counts every "
                                       "gathered row, most of which are L2 hits, so it can exceed "
                                       "1; roofline.frac uses the measured L2 gather ceiling",
                   "num_unique": num_unique, "nnz": nnz},
        "stages": stage_report,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": (round(nnz / (e2e_ms * 1e-3), 1) if e2e_ms == e2e_ms else None),
                "unit": "lookups/s",
                "ms_per_step": (round(e2e_ms, 4) if e2e_ms == e2e_ms else None),
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "protocol": "host buffers in and out every step through the public API; "
                            "copy-in / kernel / copy-out streams with two buffer sets, so "
                            "the D2H of step i overlaps step i+1; host reads num_unique "
                            "before each backward"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "wall_s_timed_region": round(t_wall, 3),
        "extras": extras,
    }
    print(json.dumps(line))


# ------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    kind_lib, kind = cpu_kind()
    threads = host_threads()
    batch = cfg["batch_size"]
    # the reference arm runs the WHOLE batch by default (about 1.5 s per step on
    # 16 host threads at C2); --ref-sample-bags bounds it for slower hosts
    sample_bags = batch if args.ref_sample_bags <= 0 else min(batch, args.ref_sample_bags)
    wl = make_host_inputs(cfg, batch)
    from cuembed_b200 import datagen
    # The table must exist at full size for the indices to be valid.
    rng = np.random.default_rng(123456)
    dt = np.float16 if cfg["dtype"] == "f16" else np.float32
    table = np.empty((cfg["num_categories"], cfg["embed_width"]), dt)
    step_rows = 1 << 20
    for r0 in range(0, cfg["num_categories"], step_rows):
        r1 = min(cfg["num_categories"], r0 + step_rows)
        table[r0:r1] = (rng.random((r1 - r0, cfg["embed_width"]), dtype=np.float32) * 2 - 1).astype(dt)
    grad_y = datagen.make_grad_y(batch, cfg["embed_width"]).astype(dt)
    for _ in range(max(args.warmup, 0)):
        cpu_reference_pass(cfg, wl.indices, table, grad_y, sample_bags, threads, kind_lib)
    tot = 0.0
    stage = {"forward": 0.0, "transpose": 0.0, "backward": 0.0}
    s_nnz = 0
    for _ in range(args.steps):
        t, s_nnz = cpu_reference_pass(cfg, wl.indices, table, grad_y, sample_bags,
                                      threads, kind_lib)
        tot += sum(t.values())
        for k in stage:
            stage[k] += t[k]
    ms = tot / args.steps * 1e3
    value = s_nnz / (ms * 1e-3)
    sample = ((f"the whole batch ({batch} bags, {s_nnz} lookups) per step" if sample_bags == batch
               else f"first {sample_bags} of {batch} bags ({s_nnz} lookups) per step")
              + f"; fwd and bwd sliced over {threads} threads, transpose single-threaded as in "
                f"the reference")
    line = {
        "impl": "reference",
        "metric": "lookups/s (fwd + transpose + bwd, compressed grad)",
        "value": round(value, 1), "unit": "lookups/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": workload_string(args.workload, cfg),
                   "arm": ("the reference's own CPU implementation of the path "
                           + ("(oracle/_ref: its templates compiled unchanged)" if kind == "reference"
                              else "(oracle port)") + f" on {threads} host threads"),
                   "sample": sample},
        "cpu_baseline": {"value": round(value, 1), "unit": "lookups/s", "cores": threads,
                         "kind": kind, "sample": sample,
                         "stage_s": {k: round(v / args.steps, 4) for k, v in stage.items()}},
        "e2e": {"value": round(value, 1), "unit": "lookups/s",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-bags", type=int, default=8192)
    ap.add_argument("--ref-sample-bags", type=int, default=0,
                    help="--impl reference: bags per step (0 = the whole batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--separate-row-ids", action="store_true",
                    help="transpose stage as ExtractRowIdsFromFixed + Transpose (the "
                         "reference's call sequence) instead of cuembed_transpose_fixed")
    ap.add_argument("--no-graphs", action="store_true",
                    help="launch the stages directly instead of replaying CUDA graphs")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra fused-optimizer measurement")
    ap.add_argument("--trace", default=None,
                    help="N > 1: write a CUPTI per-kernel time table of 5 extra steps here")
    ap.add_argument("--select-first", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: select the rank's own lookups before the forward "
                         "(auto = off: measured slower on the weak-scaled C2 shards)")
    ap.add_argument("--known-sizes", action="store_true",
                    help="N > 1: reuse local nnz / num_unique of an earlier identical step "
                         "instead of reading them back every step")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: exchange fused over peer memory, or NCCL collectives")
    ap.add_argument("--partial-dtype", default="table", choices=["f32", "table"],
                    help="N > 1, p2p: element type of the partial sums on the wire "
                         "(table: the table's own type, i.e. 16-bit partials for fp16 / bf16 "
                         "tables; f32: exact fp32 partials)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
