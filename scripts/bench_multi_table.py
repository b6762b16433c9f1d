"""Multi-table batched forward (cuembed_forward_multi) against one launch per
table: T tables of `rows` x `width`, batch B, hotness H, power-law indices.
Prints one line per configuration.  GPU only; timing with CUDA events, L2
flushed between repetitions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cuembed_b200 as ce
from cuembed_b200 import datagen

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ms = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps


for (T, rows, width, B, H) in [(26, 100_000, 128, 2048, 1), (26, 100_000, 128, 8192, 8),
                               (32, 1_000_000, 64, 4096, 20), (8, 1_000_000, 256, 16384, 32)]:
    tables = [(torch.rand(rows, width, device=dev) * 2 - 1).half() for _ in range(T)]
    idx = [torch.from_numpy(datagen.make_workload(rows, width, B, H, alpha=1.05, seed=7 + t).indices).to(dev)
           for t in range(T)]
    out = torch.empty(B, T * width, dtype=torch.float16, device=dev)
    rets = [out[:, t * width:(t + 1) * width] for t in range(T)]
    single = [torch.empty(B, width, dtype=torch.float16, device=dev) for _ in range(T)]

    def per_table():
        for t in range(T):
            ce.EmbeddingForward(tables[t], width, idx[t], None, None, B, H, ce.CombineMode.kSum, single[t])

    def multi():
        ce.EmbeddingForwardMulti(tables, width, idx, None, None, [B] * T, [H] * T, None, rets,
                                 out_row_stride=T * width)

    a, b = timed(per_table), timed(multi)
    ok = all(torch.equal(single[t], rets[t]) for t in range(T))
    print(f"tables {T} x ({rows} x {width} f16), batch {B}, hotness {H}: one launch per table {a*1e3:.1f} us, "
          f"multi-table {b*1e3:.1f} us ({a/b:.2f}x), identical={ok}")
