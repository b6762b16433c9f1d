#!/usr/bin/env python
"""Gather-ceiling study (csrc/microbench.cu) on the C2 access streams: register
landing (LDG.128, 8 or 16 rows in flight per warp) against shared-memory landing
through per-lane bulk copies (cp.async.bulk / UBLKCP), several stage shapes.
Prints one JSON line.   python scripts/bench_gather.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import cuembed_b200 as ce
from cuembed_b200 import _lib

cfg = bench.WORKLOADS["C2"]
dev = torch.device("cuda:0")
rows, w, batch, hot = cfg["num_categories"], cfg["embed_width"], cfg["batch_size"], cfg["hotness"]
nnz = batch * hot
wl = bench.make_host_inputs(cfg, batch)
table = torch.empty(rows, w, dtype=torch.float16, device=dev).uniform_(-1, 1)
grad_y = torch.empty(batch, w, dtype=torch.float16, device=dev).uniform_(-1, 1)
indices = torch.from_numpy(wl.indices).to(dev)
row_ids = torch.empty(nnz, dtype=torch.int32, device=dev)
t_idx, t_sid = torch.empty_like(indices), torch.empty_like(indices)
work = torch.empty(ce.Transpose(row_ids, indices, None, nnz, None, None, None, None),
                   dtype=torch.uint8, device=dev)
ce.ExtractRowIdsFromFixed(batch, hot, row_ids)
ce.Transpose(row_ids, indices, None, nnz, t_idx, t_sid, None, work)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
sink = torch.zeros(1, dtype=torch.int32, device=dev)
lib = _lib.load()
g = torch.Generator(device=dev); g.manual_seed(42)
rnd = torch.randint(0, batch, (nnz,), generator=g, device=dev, dtype=torch.int32)
stream = torch.cuda.current_stream()

def run(fn, flush_first, reps=5):
    best = None
    for it in range(reps + 1):
        if flush_first:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); rc = fn(); e1.record(stream); torch.cuda.synchronize()
        if rc != 0:
            return {"error": rc}
        ms = e0.elapsed_time(e1)
        if it > 0:
            best = ms if best is None else min(best, ms)
    return {"ms": round(best, 4), "GBps": round(nnz * 512 / (best * 1e-3) / 1e9, 1)}

streams = {"l2_random": (grad_y, rnd, False), "bwd_stream": (grad_y, t_sid, True),
           "fwd_stream": (table, indices, True)}
out = {}
for name, (buf, idx, fl) in streams.items():
    res = {}
    for mode, label in ((0, "ldg_8x32warps"), (2, "ldg_16x16warps"), (1, "ldg_L1_no_allocate"),
                        (3, "ldg_L2_evict_last_rows_evict_first_indices")):
        res[label] = run(lambda: lib.cuembed_microbench_gather(
            buf.data_ptr(), 512, idx.data_ptr(), nnz, mode, sink.data_ptr(), stream.cuda_stream), fl)
    for v, label in ((0, "bulk_32x2x6"), (1, "bulk_16x3x8"), (2, "bulk_32x1x12"),
                     (3, "bulk_16x2x12"), (4, "bulk_8x4x12")):
        res[label] = run(lambda: lib.cuembed_microbench_gather_bulk(
            buf.data_ptr(), 512, idx.data_ptr(), nnz, v, sink.data_ptr(), stream.cuda_stream), fl)
    for v, label in ((0, "async_2x4x6"), (1, "async_3x4x4"), (2, "async_2x4x4"),
                     (3, "async_4x4x3"), (4, "async_2x8x3")):
        res[label] = run(lambda: lib.cuembed_microbench_gather_async(
            buf.data_ptr(), 512, idx.data_ptr(), nnz, v, sink.data_ptr(), stream.cuda_stream), fl)
    out[name] = res
print(json.dumps(out))
