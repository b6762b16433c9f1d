"""Test-only local ops for cuembed_b200.sharded built on the CPU oracle, so
that the host logic (row partition, reduce-scatter / all-gather plumbing, mean
by global bag length) runs under gloo on CPU.  Not a product path."""
from __future__ import annotations

import numpy as np
import torch

from cuembed_b200.api import CombineMode
from oracle import cpu_lib


class OracleLocalOps:
    def __init__(self):
        self.lib = cpu_lib.CpuLib("oracle")

    def shard_select(self, indices, offsets, weights, batch, num_hots, lo, hi):
        idx = indices.numpy()
        off = offsets.numpy() if offsets is not None else np.arange(batch + 1) * num_hots
        keep = (idx >= lo) & (idx < hi)
        counts = np.array([keep[off[b]:off[b + 1]].sum() for b in range(batch)], np.int64)
        local_offsets = np.zeros(batch + 1, np.int32)
        np.cumsum(counts, out=local_offsets[1:])
        nnz_total = int(off[batch])
        sel = keep[:nnz_total]
        local_indices = (idx[:nnz_total][sel] - lo).astype(idx.dtype)
        pad = max(1, idx.shape[0]) - local_indices.shape[0]
        local_indices = np.concatenate([local_indices, np.zeros(pad, idx.dtype)])
        lw = None
        if weights is not None:
            w = weights.numpy()[:nnz_total][sel]
            lw = torch.from_numpy(np.concatenate([w, np.zeros(pad, w.dtype)]))
        return torch.from_numpy(local_offsets), torch.from_numpy(local_indices), lw

    def pool_partial(self, table, local_indices, local_offsets, local_weights, batch):
        w = local_weights.numpy() if local_weights is not None else None
        out = self.lib.forward(table.numpy(), local_indices.numpy(), local_offsets.numpy(),
                               w, batch, 0, cpu_lib.SUM, embed_width=table.shape[1],
                               out_dt=cpu_lib.F32)
        return torch.from_numpy(out)

    def finalize(self, partial, mode, offsets, num_hots, sample0, weights, out_dtype):
        n, width = partial.shape
        out = partial.clone()
        if mode == CombineMode.kMean:
            for s in range(n):
                if offsets is not None:
                    a, b = int(offsets[sample0 + s]), int(offsets[sample0 + s + 1])
                else:
                    a, b = (sample0 + s) * num_hots, (sample0 + s + 1) * num_hots
                if weights is not None:
                    denom = np.float32(0)
                    for j in range(a, b):
                        denom = np.float32(denom + np.float32(weights[j]))
                else:
                    denom = np.float32(b - a)
                out[s] = 0 if denom == 0 else out[s] * np.float32(np.float32(1.0) / denom)
        return out.to(out_dtype)

    def local_backward(self, grad_y, local_offsets, local_indices, local_weights,
                       batch, local_nnz, num_local_rows, compressed):
        idx = np.ascontiguousarray(local_indices.numpy()[:local_nnz])
        w = np.ascontiguousarray(local_weights.numpy()[:local_nnz]) \
            if local_weights is not None else None
        rows = self.lib.extract_row_ids_csr(local_offsets.numpy(), batch, idx.dtype)
        t_idx, t_sid, t_w = self.lib.transpose(rows, idx, w)
        if local_nnz == 0:
            n = 0 if compressed else num_local_rows
            return torch.zeros(n, grad_y.shape[1], dtype=grad_y.dtype), \
                (torch.zeros(0, dtype=local_indices.dtype) if compressed else None)
        remapped = self.lib.compressed_grad_indices(t_idx) if compressed else None
        n = int(remapped[-1]) + 1 if compressed else num_local_rows
        g, inv = self.lib.backward(grad_y.numpy(), grad_y.shape[1], n, t_idx, t_sid,
                                   remapped, t_w)
        return torch.from_numpy(g), (torch.from_numpy(inv) if inv is not None else None)
