#!/usr/bin/env python
"""Times the transpose stage alone (row ids + sort + compressed remap) on
uniform random keys; used for tuning experiments."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuembed_b200 as ce

nnz = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
idt = torch.int64 if (len(sys.argv) > 3 and sys.argv[3] == "i64") else torch.int32
dev = "cuda:0"
g = torch.Generator(device=dev); g.manual_seed(1)
idx = torch.randint(0, rows, (nnz,), generator=g, device=dev).to(idt)
row_ids = torch.empty(nnz, dtype=idt, device=dev)
t_idx = torch.empty_like(idx); t_sid = torch.empty_like(idx); rem = torch.empty_like(idx)
lw = max(ce.Transpose(row_ids, idx, None, nnz, None, None, None, None),
         ce.ComputeCompressedGradIndices(idx, nnz, None, None))
work = torch.empty(lw, dtype=torch.uint8, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def step():
    ce.ExtractRowIdsFromFixed(nnz // 64, 64, row_ids)
    ce.Transpose(row_ids, idx, None, nnz, t_idx, t_sid, None, work)
    ce.ComputeCompressedGradIndices(t_idx, nnz, rem, work)
for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.fill_(1)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ok = bool((t_idx[1:] >= t_idx[:-1]).all().item())
print(f"transpose nnz={nnz} rows={rows} {idt}: {sum(ts)/len(ts):.4f} ms (min {min(ts):.4f}) sorted={ok} env="
      + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("CUEMBED_")))
