#!/usr/bin/env python
"""Same-box GPU baseline: times the reference's own kernels (compiled unchanged
from /root/reference into oracle/_ref/libcuembed_refgpu.so) and this library on
the identical C2 inputs with the identical protocol (L2 flush before every
stage, CUDA events), and cross-checks the results.  Bench infrastructure only.

    python scripts/bench_ref_gpu.py [workload] [steps]
"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import cuembed_b200 as ce
from cuembed_b200 import _lib

workload = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = bench.WORKLOADS[workload]
dev = torch.device("cuda:0")
tdt = {"f16": torch.float16, "f32": torch.float32}[cfg["dtype"]]
dtc = {"f16": 1, "f32": 0}[cfg["dtype"]]
rows, w, batch, hot = cfg["num_categories"], cfg["embed_width"], cfg["batch_size"], cfg["hotness"]
nnz = batch * hot
wl = bench.make_host_inputs(cfg, batch)
g = torch.Generator(device=dev); g.manual_seed(123456)
table = torch.empty(rows, w, dtype=tdt, device=dev)
for r0 in range(0, rows, 1 << 20):
    r1 = min(rows, r0 + (1 << 20))
    table[r0:r1] = (torch.rand(r1 - r0, w, generator=g, device=dev) * 2 - 1).to(tdt)
g.manual_seed(654321)
grad_y = torch.randint(-10, 11, (batch, w), generator=g, device=dev).to(tdt)
indices = torch.from_numpy(wl.indices).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
vp = ctypes.c_void_p

def make_arm(lib, prefix):
    P = lambda t: vp(t.data_ptr()) if t is not None else None
    stream = lambda: vp(torch.cuda.current_stream().cuda_stream)
    st = {}
    st["out"] = torch.empty(batch, w, dtype=tdt, device=dev)
    st["row_ids"] = torch.empty(nnz, dtype=torch.int32, device=dev)
    st["t_idx"] = torch.empty(nnz, dtype=torch.int32, device=dev)
    st["t_sid"] = torch.empty(nnz, dtype=torch.int32, device=dev)
    st["rem"] = torch.empty(nnz, dtype=torch.int32, device=dev)
    lw = ctypes.c_size_t(0)
    getattr(lib, prefix + "transpose")(None, None, None, dtc, nnz, 0, None, None, None, None, ctypes.byref(lw), stream())
    lw2 = ctypes.c_size_t(0)
    getattr(lib, prefix + "compressed_grad_indices")(None, 0, nnz, None, None, ctypes.byref(lw2), stream())
    st["lw"] = ctypes.c_size_t(max(lw.value, lw2.value))
    st["work"] = torch.empty(st["lw"].value, dtype=torch.uint8, device=dev)
    def fwd():
        getattr(lib, prefix + "forward")(P(table), dtc, w, P(indices), 0, None, 0, None, batch, hot, 0, 0,
                                         P(st["out"]), dtc, stream())
    def tr():
        getattr(lib, prefix + "extract_row_ids_fixed")(batch, hot, P(st["row_ids"]), 0, stream())
        lwv = ctypes.c_size_t(st["lw"].value)
        getattr(lib, prefix + "transpose")(P(st["row_ids"]), P(indices), None, dtc, nnz, 0, P(st["t_idx"]),
                                           P(st["t_sid"]), None, P(st["work"]), ctypes.byref(lwv), stream())
        lwv = ctypes.c_size_t(st["lw"].value)
        getattr(lib, prefix + "compressed_grad_indices")(P(st["t_idx"]), 0, nnz, P(st["rem"]), P(st["work"]),
                                                         ctypes.byref(lwv), stream())
    fwd(); tr(); torch.cuda.synchronize()
    nu = int(st["rem"][-1].item()) + 1
    st["nu"] = nu
    st["grad"] = torch.zeros(nu, w, dtype=tdt, device=dev)
    st["inv"] = torch.empty(nu, dtype=torch.int32, device=dev)
    def bwd():
        getattr(lib, prefix + "backward")(P(grad_y), dtc, w, nu, nnz, 0, P(st["t_idx"]), P(st["t_sid"]),
                                          P(st["rem"]), None, 1, P(st["grad"]), P(st["inv"]), stream())
    return st, [("forward", fwd), ("transpose", tr), ("backward", bwd)]

def time_arm(stages, graph=False):
    """graph=True: every stage captured once into a CUDA graph and replayed,
    for both arms alike (bench.py's default launch mode)."""
    res = {}
    for name, fn in stages:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if graph:
            try:
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    fn()
                fn = g_.replay
                fn()
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                res[name] = float("nan")
                print(f"graph capture failed for {name}: {type(e).__name__}: {e}", file=sys.stderr)
                torch.cuda.synchronize()
                continue
        tot = 0.0
        for _ in range(steps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[name] = tot / steps
    return res

ours = _lib.load()
ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcuembed_refgpu.so"))
for lib, prefix in ((ours, "cuembed_"), (ref, "refgpu_")):
    for fn in ("forward", "extract_row_ids_fixed", "transpose", "compressed_grad_indices", "backward"):
        getattr(lib, prefix + fn).restype = ctypes.c_int

st_o, stages_o = make_arm(ours, "cuembed_")
st_r, stages_r = make_arm(ref, "refgpu_")
t_o = time_arm(stages_o)
t_r = time_arm(stages_r)
tg_o = time_arm(stages_o, graph=True)
tg_r = time_arm(stages_r, graph=True)
# cross-check (the reference GPU backward accumulates in fp16 with atomics:
# compare in value with a tolerance; the backward buffers are re-zeroed first)
st_o["grad"].zero_(); st_r["grad"].zero_()
stages_o[2][1](); stages_r[2][1](); torch.cuda.synchronize()
chk = {
    "forward_equal": bool(torch.equal(st_o["out"], st_r["out"])),
    "transpose_indices_equal": bool(torch.equal(st_o["t_idx"], st_r["t_idx"])),
    "transpose_sample_ids_equal": bool(torch.equal(st_o["t_sid"], st_r["t_sid"])),
    "remapped_equal": bool(torch.equal(st_o["rem"], st_r["rem"])),
    "inverse_mapping_equal": bool(torch.equal(st_o["inv"], st_r["inv"])),
    "backward_max_abs_diff": float((st_o["grad"].float() - st_r["grad"].float()).abs().max().item()),
    "backward_frac_equal": float((st_o["grad"] == st_r["grad"]).float().mean().item()),
}
out = {"workload": workload, "steps": steps, "nnz": nnz,
       "ours_ms": {k: round(v, 4) for k, v in t_o.items()},
       "reference_gpu_ms": {k: round(v, 4) for k, v in t_r.items()},
       "speedup": {k: round(t_r[k] / t_o[k], 3) for k in t_o},
       "total_ms": {"ours": round(sum(t_o.values()), 4), "reference_gpu": round(sum(t_r.values()), 4)},
       "graph_replay": {"ours_ms": {k: round(v, 4) for k, v in tg_o.items()},
                        "reference_gpu_ms": {k: round(v, 4) for k, v in tg_r.items()},
                        "total_ms": {"ours": round(sum(tg_o.values()), 4),
                                     "reference_gpu": round(sum(tg_r.values()), 4)}},
       "cross_check": chk}
print(json.dumps(out))
