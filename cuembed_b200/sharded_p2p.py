"""Row-sharded embedding with the exchange fused into the kernels over NVLink /
NVSwitch peer memory (csrc/sharded_p2p.cu); same interface as
cuembed_b200.sharded.RowShardedEmbedding, no collective call on the data path.

forward   cuembed_shard_pool_push  (gather + pool + store into the bag owner's
          slot over NVLink) -> cuembed_shard_signal -> cuembed_shard_reduce_finalize
          (wait for all ranks, sum the slots in rank order, mean by the global
          bag length, cast).
          concat: cuembed_shard_concat_push writes every looked-up row to its
          final place in the bag owner's output (all-to-all made of stores).
backward  the grad_y slice is pushed to every rank by the copy engines on a
          side stream (cuembed_shard_allgather_push + signal) WHILE the main
          stream selects and sorts the rank's own lookups; cuembed_shard_wait;
          ordinary deterministic backward on the owner.  Gradients never leave
          the owning shard.

Exchange buffers are double-buffered by epoch parity (see the protocol note in
csrc/sharded_p2p.cu); every rank must call forward / backward the same number
of times, as with any collective.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import _lib, peer
from .api import CombineMode, CuEmbedError, _check, _dev, _dt, _it, _stream
from .sharded import CudaLocalOps, row_range

CH_FORWARD, CH_GRAD, CH_CONCAT = 0, 1, 2


@dataclass
class PeerForwardContext:
    indices: torch.Tensor
    offsets: Optional[torch.Tensor]
    weights: Optional[torch.Tensor]
    counts: Optional[torch.Tensor]
    batch: int
    num_hots: int
    mode: CombineMode
    coo: Optional[tuple] = None
    local_nnz: int = -1
    selected: Optional[tuple] = None   # (local_offsets, local_indices, sample_ids, local_weights)
    # the number of owned lookups, summed from the pool kernel's per-bag counts and
    # copied to the host on the side stream while the exchange runs
    nnz_host: Optional[torch.Tensor] = None
    nnz_event: Optional[torch.cuda.Event] = None


class PeerShardedEmbedding:
    """One rank's shard; exchange through peer memory.

    `buffers(kind, nbytes)` returns this rank's view of a symmetric allocation
    (peer.PeerBuffer by default; tests pass peer.LocalPeerGroup views together
    with explicit `rank` / `world` to run several virtual ranks in one process).
    partial_dtype: element type of the partial sums on the wire.  None (default)
    = the table's own type: fp32 partials for fp32 tables (bit-exact), 16-bit
    partials for fp16 / bf16 tables (half the NVLink bytes; every partial sum is
    rounded once to the table type before the rank-ordered fp32 reduction, so
    the result differs from the fp32-partial result by at most world roundings
    of the 16-bit type -- the same order as the final rounding of the output).
    torch.float32 forces exact fp32 partials for any table.

    A peer that never signals makes the waiting kernel give up after the peer
    timeout (default 60 s, set_peer_timeout_ms): the affected output / gather
    slice is filled with NaN and the next call on this object raises.
    """

    def __init__(self, local_table: torch.Tensor, num_rows: int,
                 group: Optional[dist.ProcessGroup] = None,
                 partial_dtype: Optional[torch.dtype] = None,
                 rank: Optional[int] = None, world: Optional[int] = None,
                 buffers: Optional[Callable] = None,
                 select_first: Optional[bool] = None):
        if not local_table.is_cuda:
            raise CuEmbedError("PeerShardedEmbedding needs CUDA tensors: there is no CPU path")
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.num_rows = num_rows
        self.lo, self.hi = row_range(num_rows, self.world, self.rank)
        if local_table.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns rows [{self.lo}, {self.hi}) but the "
                             f"local table has {local_table.shape[0]} rows")
        self.table = local_table
        self.device = local_table.device
        self.partial_dtype = local_table.dtype if partial_dtype is None else partial_dtype
        self.ops = CudaLocalOps()
        # opt-in: select the rank's own lookups BEFORE the forward (one scan of
        # the replicated index list for forward and backward).  Measured at
        # N = 8: 5 % faster on C5 (width 128, one global table), 10 % slower on
        # the weak-scaled C2 shards -- so the default lets the pool-and-push
        # kernel filter the list itself.
        self.select_first = bool(select_first) if select_first is not None else False
        self._lib = _lib.load()
        self._buffers = buffers if buffers is not None else \
            (lambda kind, nbytes: peer.PeerBuffer(nbytes, group, self.device))
        self._bufs = {}
        self._epoch = {CH_FORWARD: 0, CH_GRAD: 0, CH_CONCAT: 0}
        self._side = torch.cuda.Stream(device=self.device)
        self._flags = self._buffers("flags", peer.FLAG_BYTES)
        self._flag_ptrs = peer.ptr_array(self._flags.ptrs)
        # one step in flight per channel: the exchange buffers are two deep (epoch
        # parity), which only protects begin(e) / finish(e) / begin(e+1) / ...
        self._pending = {CH_FORWARD: False, CH_GRAD: False, CH_CONCAT: False}
        # asynchronous read-back of the wait status (checked one call later, so
        # the data path never synchronises with the host)
        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._status_event = None
        self._nnz_host = [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(2)]

    # ------------------------------------------------------------ plumbing
    def _buffer(self, kind: str, nbytes: int):
        buf = self._bufs.get(kind)
        if buf is None or buf.nbytes < nbytes:
            if buf is not None:
                buf.close()
            buf = self._buffers(kind, nbytes)
            self._bufs[kind] = buf
        return buf

    def _signal(self, channel: int, epoch: int, stream) -> None:
        _check(self._lib.cuembed_shard_signal(self._flag_ptrs, self.world, self.rank,
                                              channel, epoch, _stream(stream)))

    def _status_tensor(self) -> torch.Tensor:
        word = peer.CHANNELS * peer.MAX_WORLD
        return self._flags.tensor(4 * word, (1,), torch.int32)

    def status(self) -> int:
        """0, or 1 + the rank a wait gave up on (synchronises)."""
        return int(self._status_tensor().item())

    def set_peer_timeout_ms(self, timeout_ms: int) -> None:
        """Process-wide: how long a wait spins before it gives up."""
        _check(self._lib.cuembed_shard_set_timeout_ms(int(timeout_ms)))

    def _begin(self, channel: int) -> None:
        self._raise_on_timeout()
        if self._pending[channel]:
            raise CuEmbedError(
                "a step of this channel is still in flight: call *_finish before the "
                "next *_begin (the peer exchange buffers are two epochs deep)")
        self._pending[channel] = True

    def _finish(self, channel: int, stream=None) -> None:
        """After the waiting kernel of a step: queue the status read-back."""
        self._pending[channel] = False
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        if isinstance(s, torch.cuda.Stream):
            with torch.cuda.stream(s):
                self._status_host.copy_(self._status_tensor(), non_blocking=True)
                self._status_event = torch.cuda.Event()
                self._status_event.record(s)

    def _raise_on_timeout(self) -> None:
        ev = self._status_event
        if ev is not None and ev.query():
            code = int(self._status_host.item())
            if code != 0:
                raise CuEmbedError(
                    f"rank {self.rank}: a peer wait timed out waiting for rank {code - 1}; "
                    f"the affected results were filled with NaN")

    def close(self) -> None:
        for buf in self._bufs.values():
            buf.close()
        self._bufs = {}
        self._flags.close()

    # ------------------------------------------------------------- forward
    def forward(self, indices, offsets, weights, batch_size: int, num_hots: int,
                mode: CombineMode = CombineMode.kSum,
                out_dtype: Optional[torch.dtype] = None, stream=None,
                local_nnz: Optional[int] = None):
        """Pooled lookup of the GLOBAL batch (replicated indices); returns this
        rank's slice [batch/world, width] (concat: [batch/world * hots, width])
        and the context for backward."""
        return self.forward_finish(self.forward_begin(
            indices, offsets, weights, batch_size, num_hots, mode, out_dtype, stream,
            local_nnz=local_nnz))

    def forward_begin(self, indices, offsets, weights, batch_size: int, num_hots: int,
                      mode: CombineMode = CombineMode.kSum,
                      out_dtype: Optional[torch.dtype] = None, stream=None,
                      local_nnz: Optional[int] = None):
        """Push phase (never waits for another rank).  Returns a pending handle
        for forward_finish.

        With `select_first` (constructor option, off by default) the rank first selects its
        own lookups from the replicated index list (one pass, also what the
        backward needs) and the pool-and-push kernel then walks only those; with
        fewer ranks the push kernel filters the replicated list itself.
        `local_nnz`: the number of lookups this rank owns, if the caller knows
        it (saves the host read that sizes the selection)."""
        if batch_size % self.world != 0:
            raise ValueError("batch_size must be divisible by the number of ranks")
        if weights is not None and weights.dtype != self.table.dtype:
            raise CuEmbedError("weights must have the element type of the table")
        if mode == CombineMode.kConcat:
            return self._concat_begin(indices, offsets, weights, batch_size, num_hots,
                                      stream)
        lib = self._lib
        self._begin(CH_FORWARD)
        width = self.table.shape[1]
        per = batch_size // self.world
        out_dtype = self.table.dtype if out_dtype is None else out_dtype
        psize = torch.empty(0, dtype=self.partial_dtype).element_size()
        one = self.world * per * width * psize
        buf = self._buffer("slots", 2 * one)
        self._epoch[CH_FORWARD] += 1
        epoch = self._epoch[CH_FORWARD]
        base = (epoch & 1) * one
        slot_ptrs = peer.ptr_array(buf.ptrs, base + self.rank * per * width * psize)
        counts = torch.empty(batch_size, dtype=torch.int32, device=self.device)
        push_idx, push_off, push_w, push_hots = indices, offsets, weights, num_hots
        lo, hi = self.lo, self.hi
        selected = None
        if self.select_first:
            # The push kernel at N ranks would scan N x its share of indices; the
            # selection scans them once for forward AND backward.
            concat_w = weights
            selected = self.ops.shard_select_coo(indices, offsets, concat_w, batch_size,
                                                 num_hots, self.lo, self.hi, counts=None,
                                                 nnz_cap=local_nnz)
            push_off, push_idx, _, push_w = selected
            push_hots, lo, hi = 0, 0, self.hi - self.lo
        _check(lib.cuembed_shard_pool_push(
            _dev(self.table, "table"), _dt(self.table), width,
            _dev(push_idx, "indices"), _it(push_idx), _dev(push_off, "offsets"),
            _it(push_off) if push_off is not None else 0, _dev(push_w, "weights"),
            batch_size, push_hots, lo, hi, slot_ptrs, self.world, self.rank,
            _dt(torch.empty(0, dtype=self.partial_dtype)), _dev(counts, "counts"),
            _stream(stream)))
        self._signal(CH_FORWARD, epoch, stream)
        ctx = PeerForwardContext(indices, offsets, weights, counts, batch_size,
                                 num_hots, mode)
        ctx.selected = selected
        if local_nnz is not None:
            ctx.local_nnz = int(local_nnz)
        elif selected is None and (stream is None or isinstance(stream, torch.cuda.Stream)):
            # The backward's sort is sized by the number of lookups this rank owns.
            # The pool kernel has just counted them per bag: sum the counts and copy
            # the total to the host on the side stream, under the exchange, so that
            # prepare_backward finds it there instead of stalling the main stream
            # on a read-back after the select.
            main = stream if stream is not None else torch.cuda.current_stream(self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                host = self._nnz_host[epoch & 1]
                host.copy_(counts.sum(dtype=torch.int64).view(1), non_blocking=True)
                ctx.nnz_host = host
                ctx.nnz_event = torch.cuda.Event()
                ctx.nnz_event.record(self._side)
            counts.record_stream(self._side)
        return ("pool", ctx, buf, base, epoch, out_dtype, stream)

    def forward_finish(self, pending):
        """Wait for every rank's pushes, reduce in rank order, epilogue."""
        if pending[0] == "concat":
            return self._concat_finish(pending)
        _, ctx, buf, base, epoch, out_dtype, stream = pending
        lib = self._lib
        width = self.table.shape[1]
        per = ctx.batch // self.world
        offsets, weights, num_hots, mode = ctx.offsets, ctx.weights, ctx.num_hots, ctx.mode
        out = torch.empty(per, width, dtype=out_dtype, device=self.device)
        _check(lib.cuembed_shard_reduce_finalize(
            buf.local + base, _dt(torch.empty(0, dtype=self.partial_dtype)), self.world,
            self._flags.local, CH_FORWARD, epoch, per, width, int(mode),
            _dev(offsets, "offsets"), _it(offsets) if offsets is not None else 0,
            num_hots, self.rank * per, _dev(weights, "weights"),
            _dt(weights) if weights is not None else 0, _dev(out, "out"), _dt(out),
            _stream(stream)))
        self._finish(CH_FORWARD, stream)
        return out, ctx

    def _concat_begin(self, indices, offsets, weights, batch_size, num_hots, stream):
        if weights is not None:
            raise CuEmbedError("Check failed: weights == nullptr || mode != CombineMode::kConcat")
        if offsets is not None or num_hots <= 0:
            raise CuEmbedError("Check failed: offsets == nullptr || mode != CombineMode::kConcat")
        lib = self._lib
        self._begin(CH_CONCAT)
        width = self.table.shape[1]
        per = batch_size // self.world
        one = per * num_hots * width * self.table.element_size()
        buf = self._buffer("concat", 2 * one)
        self._epoch[CH_CONCAT] += 1
        epoch = self._epoch[CH_CONCAT]
        base = (epoch & 1) * one
        _check(lib.cuembed_shard_concat_push(
            _dev(self.table, "table"), _dt(self.table), width, _dev(indices, "indices"),
            _it(indices), batch_size, num_hots, self.lo, self.hi,
            peer.ptr_array(buf.ptrs, base), self.world, self.rank, _stream(stream)))
        self._signal(CH_CONCAT, epoch, stream)
        ctx = PeerForwardContext(indices, None, None, None, batch_size, num_hots,
                                 CombineMode.kConcat)
        return ("concat", ctx, buf, base, epoch, None, stream)

    def _concat_finish(self, pending):
        _, ctx, buf, base, epoch, _, stream = pending
        lib = self._lib
        width = self.table.shape[1]
        per = ctx.batch // self.world
        num_hots = ctx.num_hots
        one = per * num_hots * width * self.table.element_size()
        _check(lib.cuembed_shard_wait(self._flags.local, self.world, CH_CONCAT, epoch,
                                      buf.local + base, one // self.world,
                                      _stream(stream)))
        self._finish(CH_CONCAT, stream)
        # valid until the next-but-one concat forward (double buffer)
        out = buf.tensor(base, (per * num_hots, width), self.table.dtype)
        return out, ctx

    # ------------------------------------------------------------ backward
    def prepare_backward(self, ctx: PeerForwardContext, compressed: bool = True,
                         local_nnz: Optional[int] = None) -> None:
        """Select + sort the rank's own lookups (independent of grad_y; called by
        backward if the caller has not done it earlier).  `local_nnz`: the
        number of lookups this rank owns if the caller already knows it (saves
        one host read)."""
        if ctx.coo is not None:
            return
        concat = ctx.mode == CombineMode.kConcat
        weights = ctx.weights
        if concat:
            # the sample id of a concat lookup is its position in the global
            # index list: carry it through the select as a 4-byte payload
            if ctx.indices.numel() >= 2 ** 31:
                raise CuEmbedError("sharded concat backward needs nnz < 2^31")
            weights = torch.arange(ctx.indices.numel(), dtype=torch.int32,
                                   device=self.device).view(torch.float32)
        if local_nnz is None and ctx.local_nnz >= 0:
            local_nnz = ctx.local_nnz
        pending_nnz = local_nnz is None and ctx.nnz_event is not None and not concat
        if ctx.selected is not None and not concat:
            l_off, l_idx, l_sid, l_w = ctx.selected   # selected before the forward
        else:
            # With the count still on its way to the host the select is launched
            # into buffers sized for the whole index list, so that the GPU has the
            # select to run while the host waits for the number that sizes the sort.
            l_off, l_idx, l_sid, l_w = self.ops.shard_select_coo(
                ctx.indices, ctx.offsets, weights, ctx.batch, ctx.num_hots, self.lo,
                self.hi, counts=ctx.counts, nnz_cap=local_nnz)
        if pending_nnz:
            ctx.nnz_event.synchronize()        # recorded under the forward's exchange
            local_nnz = int(ctx.nnz_host.item())
        if local_nnz is not None:
            ctx.local_nnz = int(local_nnz)
        elif ctx.local_nnz < 0:
            ctx.local_nnz = int(l_off[-1].item())
        if ctx.local_nnz == 0:
            ctx.coo = ()
            return
        sample_ids = l_sid[:ctx.local_nnz]
        if concat:
            sample_ids = l_w[:ctx.local_nnz].view(torch.int32).to(l_idx.dtype)
            l_w = None
        ctx.coo = self.ops.local_transpose(l_off, l_idx, l_w, ctx.batch, ctx.local_nnz,
                                           compressed, sample_ids=sample_ids)

    def backward(self, grad_out_slice: torch.Tensor, ctx: PeerForwardContext,
                 compressed: bool = True):
        """grad_out_slice: this rank's slice of dL/dout.  Returns (grad, rows):
        gradient rows for this shard and, if compressed, the GLOBAL table row of
        each gradient row."""
        return self.backward_finish(self.backward_begin(grad_out_slice, ctx, compressed))

    def backward_begin(self, grad_out_slice: torch.Tensor, ctx: PeerForwardContext,
                       compressed: bool = True, local_nnz: Optional[int] = None):
        """Push phase: copy engines send the slice to every rank (side stream)
        while the main stream selects and sorts; never waits for another rank."""
        lib = self._lib
        self._begin(CH_GRAD)
        grad_out_slice = grad_out_slice.contiguous()
        width = grad_out_slice.shape[1]
        n_slice = grad_out_slice.shape[0]
        es = grad_out_slice.element_size()
        one = self.world * n_slice * width * es
        buf = self._buffer("gather", 2 * one)
        self._epoch[CH_GRAD] += 1
        epoch = self._epoch[CH_GRAD]
        base = (epoch & 1) * one
        main = torch.cuda.current_stream(self.device)
        # copy engines push the slice to every rank while the SMs sort
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            _check(lib.cuembed_shard_allgather_push(
                _dev(grad_out_slice, "grad_out_slice"), n_slice * width * es,
                peer.ptr_array(buf.ptrs, base), self.world, self.rank,
                self._side.cuda_stream))
            self._signal(CH_GRAD, epoch, self._side)
        grad_out_slice.record_stream(self._side)
        self.prepare_backward(ctx, compressed, local_nnz)
        return (grad_out_slice, ctx, compressed, buf, base, epoch)

    def backward_finish(self, pending, grad=None, inverse=None):
        """`grad` [num_unique or shard rows, width] / `inverse` [num_unique]:
        optional preallocated outputs (then nothing is read back to size them)."""
        grad_out_slice, ctx, compressed, buf, base, epoch = pending
        lib = self._lib
        width = grad_out_slice.shape[1]
        n_slice = grad_out_slice.shape[0]
        main = torch.cuda.current_stream(self.device)
        _check(lib.cuembed_shard_wait(self._flags.local, self.world, CH_GRAD, epoch,
                                      buf.local + base,
                                      n_slice * width * grad_out_slice.element_size(),
                                      main.cuda_stream))
        main.wait_stream(self._side)  # the local slice was copied on the side stream
        self._finish(CH_GRAD, main)
        if ctx.local_nnz == 0:
            rows = 0 if compressed else self.hi - self.lo
            idt = ctx.indices.dtype
            return (torch.zeros(rows, width, dtype=grad_out_slice.dtype, device=self.device),
                    torch.empty(0, dtype=idt, device=self.device) if compressed else None)
        full = buf.tensor(base, (self.world * n_slice, width), grad_out_slice.dtype)
        grad, inv = self.ops.local_backward_coo(full, ctx.coo, ctx.local_nnz,
                                                self.hi - self.lo, grad, inverse)
        rows = inv + self.lo if inv is not None else None
        return grad, rows
