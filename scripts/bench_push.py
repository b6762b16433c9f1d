#!/usr/bin/env python
"""ShardPoolPushKernel alone, on ONE GPU, as rank 0 of a VIRTUAL world of W ranks
(peer.LocalPeerGroup: every slot buffer lives on this GPU, so the stores do not
cross NVLink -- what remains is the kernel's own structure: scan W x its share of
the replicated index list, gather the owned rows, store one partial row per bag).
Workload: the weak-scaled C2 shard of the N > 1 bench (10 M x 256 fp16 rows per
rank, global batch W x 65 536, hotness 64, alpha 1.15).
    python scripts/bench_push.py [W ...]      prints one JSON line per W"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cuembed_b200 as ce
from cuembed_b200 import _lib, peer
from cuembed_b200.api import _dev, _dt, _it, _check
from benchmarks.sharded_bench import unique_bags_torch

dev = torch.device("cuda:0")
lib = _lib.load()
worlds = [int(a) for a in sys.argv[1:]] or [2, 8]
shard_rows, w, hot, per = 10_000_000, 256, 64, 65536
table = torch.empty(shard_rows, w, dtype=torch.float16, device=dev).uniform_(-1, 1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for W in worlds:
    rows, batch = shard_rows * W, per * W
    g = torch.Generator(device=dev); g.manual_seed(1234)
    indices = unique_bags_torch(g, batch, hot, rows, 1.15, dev).view(-1).to(torch.int32)
    group = peer.LocalPeerGroup(W, dev)
    one = W * per * w * 2
    views = group.alloc(one)
    slot_ptrs = peer.ptr_array(views[0].ptrs, 0)     # rank 0's slot [0] in every owner's buffer
    counts = torch.empty(batch, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    def push():
        _check(lib.cuembed_shard_pool_push(
            table.data_ptr(), 1, w, indices.data_ptr(), 0, None, 0, None, batch, hot,
            0, shard_rows, slot_ptrs, W, 0, 1, counts.data_ptr(), stream.cuda_stream))
    times = []
    for it in range(6):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); push(); e1.record(stream); torch.cuda.synchronize()
        if it: times.append(e0.elapsed_time(e1))
    kept = int(counts.sum().item())
    # the backward's select (counts from the push kernel): scan + fill of the local COO
    from cuembed_b200.sharded import CudaLocalOps
    ops = CudaLocalOps()
    sel = []
    for it in range(6):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ops.shard_select_coo(indices, None, None, batch, hot, 0, shard_rows, counts=counts,
                             nnz_cap=kept)
        e1.record(stream); torch.cuda.synchronize()
        if it: sel.append(e0.elapsed_time(e1))
    print(json.dumps({"world": W, "global_batch": batch, "lookups_scanned": batch * hot,
                      "lookups_owned": kept, "push_kernel_ms": round(min(times), 4),
                      "median_ms": round(sorted(times)[len(times) // 2], 4),
                      "select_coo_ms": round(min(sel), 4)}))
    group.close()
    del indices, counts
