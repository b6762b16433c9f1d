"""Row-sharded multi-GPU embedding (one process per GPU, torch.distributed).

New functionality relative to the reference, which is single-table single-GPU
(README.md:110); specified by BASELINE.json `north_star` and SURVEY.md 8(e):

  * the table is split into contiguous row ranges, one per rank;
  * forward: every rank pools the lookups that hit ITS rows into an fp32
    partial [batch, width]; an NCCL reduce-scatter over NVLink / NVSwitch sums
    the partials and leaves each rank with its slice of the batch; mean divides
    by the GLOBAL bag length after the reduction;
  * backward: all-gather of grad_y (the adjoint of the reduce-scatter), then
    the ordinary transpose + backward on the rank's own lookups -- gradients
    never leave the owning shard.

Two transports for the exchange step:

  * "p2p" (default on CUDA when the ranks can map each other's memory): the
    exchange is fused into the kernels over NVLink / NVSwitch peer memory
    (csrc/sharded_p2p.cu) -- the pooling kernel stores every partial row
    straight into the bag owner's slot while it gathers, the owner sums the
    slots in rank order (deterministic), grad_y slices are pushed by the copy
    engines while the local transpose runs; no collective call on the data path;
  * "nccl": reduce-scatter / all-gather through torch.distributed (the
    baseline, and what the gloo host-logic tests exercise on CPU).

The local compute is the single-GPU library (C ABI) on the rank's compact local
CSR produced by cuembed_shard_select.  `ops` is injectable so that the host
logic (partitioning, collectives, epilogue rules) can be exercised with gloo on
CPU in tests; the default and only product implementation is CUDA.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .api import CombineMode


def row_range(num_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row ownership: owner(row) = row // ceil(num_rows / world)."""
    per = (num_rows + world - 1) // world
    lo = min(num_rows, rank * per)
    hi = min(num_rows, lo + per)
    return lo, hi


class CudaLocalOps:
    """Local stages on the rank's GPU through libcuembed_b200.so."""

    def __init__(self):
        from . import api
        self.api = api
        self._work = {}

    def _scratch(self, key, nbytes, device):
        buf = self._work.get(key)
        if buf is None or buf.numel() < nbytes or buf.device != device:
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self._work[key] = buf
        return buf

    def shard_select(self, indices, offsets, weights, batch, num_hots, lo, hi):
        api = self.api
        dev = indices.device
        nnz_cap = indices.numel()
        local_offsets = torch.empty(batch + 1, dtype=torch.int32, device=dev)
        local_indices = torch.empty(max(nnz_cap, 1), dtype=indices.dtype, device=dev)
        local_weights = torch.empty(max(nnz_cap, 1), dtype=weights.dtype, device=dev) \
            if weights is not None else None
        nbytes = api.ShardSelect(indices, offsets, weights, batch, num_hots, lo, hi,
                                 None, None, None, None)
        work = self._scratch("select", nbytes, dev)
        api.ShardSelect(indices, offsets, weights, batch, num_hots, lo, hi,
                        local_offsets, local_indices, local_weights, work)
        return local_offsets, local_indices, local_weights

    def shard_select_coo(self, indices, offsets, weights, batch, num_hots, lo, hi,
                         counts=None, nnz_cap=None):
        """Selection in COO form: (local_offsets, local_indices, sample_ids,
        local_weights); `counts` from the pooling kernel saves the counting pass,
        `nnz_cap` (the number of owned lookups, if known) sizes the outputs."""
        api = self.api
        dev = indices.device
        cap = indices.numel() if nnz_cap is None else nnz_cap
        local_offsets = torch.empty(batch + 1, dtype=torch.int32, device=dev)
        local_indices = torch.empty(max(cap, 1), dtype=indices.dtype, device=dev)
        sample_ids = torch.empty(max(cap, 1), dtype=indices.dtype, device=dev)
        local_weights = torch.empty(max(cap, 1), dtype=weights.dtype, device=dev) \
            if weights is not None else None
        nbytes = api.ShardSelect(indices, offsets, weights, batch, num_hots, lo, hi,
                                 None, None, None, None)
        work = self._scratch("select", nbytes, dev)
        api.ShardSelectCoo(indices, offsets, weights, batch, num_hots, lo, hi, counts,
                           local_offsets, local_indices, sample_ids, local_weights, work)
        return local_offsets, local_indices, sample_ids, local_weights

    def pool_partial(self, table, local_indices, local_offsets, local_weights, batch):
        width = table.shape[1]
        partial = torch.empty(batch, width, dtype=torch.float32, device=table.device)
        self.api.EmbeddingForward(table, width, local_indices, local_offsets,
                                  local_weights, batch, 0, CombineMode.kSum, partial)
        return partial

    def finalize(self, partial, mode, offsets, num_hots, sample0, weights, out_dtype):
        n, width = partial.shape
        if mode == CombineMode.kSum and out_dtype == torch.float32:
            return partial
        out = torch.empty(n, width, dtype=out_dtype, device=partial.device)
        self.api.ShardFinalize(partial, n, width, mode, offsets, num_hots, sample0,
                               weights, out)
        return out

    def local_transpose(self, local_offsets, local_indices, local_weights, batch,
                        local_nnz, compressed, sample_ids=None):
        """Row ids + stable sort (+ compressed remap) of the rank's own lookups;
        independent of grad_y.  Returns (t_idx, t_sid, t_w, remapped)."""
        api = self.api
        dev = local_indices.device
        idt = local_indices.dtype
        if sample_ids is None:
            row_ids = torch.empty(local_nnz, dtype=idt, device=dev)
            api.ExtractRowIdsFromCSR(local_offsets, batch, row_ids)
        else:
            row_ids = sample_ids
        t_idx = torch.empty(local_nnz, dtype=idt, device=dev)
        t_sid = torch.empty(local_nnz, dtype=idt, device=dev)
        t_w = torch.empty(local_nnz, dtype=local_weights.dtype, device=dev) \
            if local_weights is not None else None
        idx = local_indices[:local_nnz]
        w = local_weights[:local_nnz] if local_weights is not None else None
        nbytes = max(api.Transpose(row_ids, idx, w, local_nnz, None, None, None, None),
                     api.ComputeCompressedGradIndices(idx, local_nnz, None, None))
        work = self._scratch("transpose", nbytes, dev)
        api.Transpose(row_ids, idx, w, local_nnz, t_idx, t_sid, t_w, work)
        remapped = None
        if compressed:
            remapped = torch.empty(local_nnz, dtype=idt, device=dev)
            api.ComputeCompressedGradIndices(t_idx, local_nnz, remapped, work)
        return t_idx, t_sid, t_w, remapped

    def local_backward_coo(self, grad_y, coo, local_nnz, num_local_rows,
                           grad=None, inv=None):
        """`grad` / `inv` may be preallocated by a caller that already knows the
        number of unique rows (saves the host read of remapped[-1])."""
        api = self.api
        t_idx, t_sid, t_w, remapped = coo
        dev = grad_y.device
        width = grad_y.shape[1]
        if grad is None and remapped is not None:
            # Compressed gradient of unknown size: launch into buffers sized for the
            # upper bound (every owned lookup a distinct row) and read the number of
            # unique rows back AFTER the launch, so the read-back waits under the
            # backward instead of leaving the GPU idle in front of it.  The result
            # is a view of the first `rows` rows.
            cap = max(1, min(local_nnz, num_local_rows))
            grad = torch.empty(cap, width, dtype=grad_y.dtype, device=dev)
            inv = torch.empty(cap, dtype=t_idx.dtype, device=dev)
            api.EmbeddingBackward(grad_y, width, cap, local_nnz, t_idx, t_sid, remapped,
                                  t_w, True, grad, inv)
            rows = int(remapped[-1].item()) + 1
            return grad[:rows], inv[:rows]
        if grad is None:
            grad = torch.empty(num_local_rows, width, dtype=grad_y.dtype, device=dev)
        rows = grad.shape[0]
        # every row of a compressed gradient is written: no zero-fill needed
        api.EmbeddingBackward(grad_y, width, rows, local_nnz, t_idx, t_sid, remapped,
                              t_w, remapped is not None, grad, inv)
        return grad, inv

    def local_backward(self, grad_y, local_offsets, local_indices, local_weights,
                       batch, local_nnz, num_local_rows, compressed):
        dev = grad_y.device
        width = grad_y.shape[1]
        idt = local_indices.dtype
        if local_nnz == 0:
            rows = 0 if compressed else num_local_rows
            return (torch.zeros(rows, width, dtype=grad_y.dtype, device=dev),
                    torch.empty(0, dtype=idt, device=dev) if compressed else None)
        coo = self.local_transpose(local_offsets, local_indices, local_weights, batch,
                                   local_nnz, compressed)
        return self.local_backward_coo(grad_y, coo, local_nnz, num_local_rows)


@dataclass
class ForwardContext:
    local_offsets: torch.Tensor
    local_indices: torch.Tensor
    local_weights: Optional[torch.Tensor]
    batch: int


class RowShardedEmbedding:
    """One rank's shard of a row-sharded embedding table."""

    def __init__(self, local_table: torch.Tensor, num_rows: int,
                 group: Optional[dist.ProcessGroup] = None, ops=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.num_rows = num_rows
        self.lo, self.hi = row_range(num_rows, self.world, self.rank)
        if local_table.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns rows [{self.lo}, {self.hi}) but the "
                             f"local table has {local_table.shape[0]} rows")
        self.table = local_table
        self.ops = ops if ops is not None else CudaLocalOps()

    # ------------------------------------------------------------- forward
    def forward(self, indices, offsets, weights, batch_size: int, num_hots: int,
                mode: CombineMode = CombineMode.kSum,
                out_dtype: Optional[torch.dtype] = None):
        """Pooled lookup of the GLOBAL batch (replicated indices); returns this
        rank's slice [batch/world, width] and the context for backward."""
        if mode == CombineMode.kConcat:
            raise NotImplementedError("sharded concat is not implemented yet")
        if batch_size % self.world != 0:
            raise ValueError("batch_size must be divisible by the number of ranks")
        out_dtype = self.table.dtype if out_dtype is None else out_dtype
        lo_off, lo_idx, lo_w = self.ops.shard_select(
            indices, offsets, weights, batch_size, num_hots, self.lo, self.hi)
        partial = self.ops.pool_partial(self.table, lo_idx, lo_off, lo_w, batch_size)
        per = batch_size // self.world
        mine = torch.empty(per, partial.shape[1], dtype=torch.float32,
                           device=partial.device)
        dist.reduce_scatter_tensor(mine, partial, op=dist.ReduceOp.SUM, group=self.group)
        out = self.ops.finalize(mine, mode, offsets, num_hots, self.rank * per,
                                weights, out_dtype)
        return out, ForwardContext(lo_off, lo_idx, lo_w, batch_size)

    # ------------------------------------------------------------ backward
    def backward(self, grad_out_slice: torch.Tensor, ctx: ForwardContext,
                 compressed: bool = True):
        """grad_out_slice: this rank's [batch/world, width] slice of dL/dout.
        Returns (grad, rows): gradient rows for this shard and, if compressed,
        the GLOBAL table row of each gradient row."""
        width = grad_out_slice.shape[1]
        full = torch.empty(ctx.batch, width, dtype=grad_out_slice.dtype,
                           device=grad_out_slice.device)
        dist.all_gather_into_tensor(full, grad_out_slice.contiguous(), group=self.group)
        local_nnz = int(ctx.local_offsets[-1].item())
        grad, inv = self.ops.local_backward(
            full, ctx.local_offsets, ctx.local_indices, ctx.local_weights, ctx.batch,
            local_nnz, self.hi - self.lo, compressed)
        rows = inv + self.lo if inv is not None else None
        return grad, rows
