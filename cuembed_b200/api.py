"""Host-side mirror of the cuEmbed operator interface on torch CUDA tensors.

Function names, argument order and meaning follow the reference's host API
(NVIDIA/cuEmbed, cuembed/include/embedding_lookup.cuh:245-308,423-483 and
cuembed/include/index_transforms.cuh:45-93,224-250,278-323) so that parity
tests read like the reference's own tests.  Every function enqueues sm_100a
kernels from libcuembed_b200.so on the current torch CUDA stream and returns
without synchronising.  Tensors must live on a CUDA device: there is no CPU
path (the reference's `Check failed ... abort()` becomes CuEmbedError).
"""
from __future__ import annotations

import ctypes
import enum
import os
from typing import Optional

import torch

from . import _lib


class CombineMode(enum.IntEnum):
    """cuembed::CombineMode, cuembed/include/embedding_lookup_types.cuh:29."""
    kSum = 0
    kMean = 1
    kConcat = 2


class CuEmbedError(RuntimeError):
    pass


# CUEMBED_DEBUG_CHECKS=1: EmbeddingForward validates its indices / offsets first
_DEBUG_CHECKS = os.environ.get("CUEMBED_DEBUG_CHECKS", "0") not in ("", "0")

_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_IT = {torch.int32: 0, torch.int64: 1}


def _check(rc: int) -> None:
    if rc != 0:
        msg = _lib.load().cuembed_error_string(rc).decode()
        raise CuEmbedError(f"cuembed_b200 error {rc}: {msg}")


def _dev(t: Optional[torch.Tensor], name: str) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise CuEmbedError(
            f"{name} must be a CUDA tensor: cuembed_b200 has no CPU fallback")
    if not t.is_contiguous():
        raise CuEmbedError(f"{name} must be contiguous")
    return t.data_ptr()


def _stream(stream) -> int:
    if stream is None:
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, torch.cuda.Stream):
        return stream.cuda_stream
    return int(stream)


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise CuEmbedError(f"unsupported element dtype {t.dtype}") from None


def _it(t: torch.Tensor) -> int:
    try:
        return _IT[t.dtype]
    except KeyError:
        raise CuEmbedError(f"unsupported index dtype {t.dtype}") from None


def launch_count() -> int:
    """Kernels launched by the library in this process."""
    return int(_lib.load().cuembed_launch_count())


def EmbeddingForward(params: torch.Tensor, embed_width: int,
                     indices: torch.Tensor, offsets: Optional[torch.Tensor],
                     weights: Optional[torch.Tensor], batch_size: int,
                     num_hots: int, mode: CombineMode, ret: torch.Tensor,
                     fp16_math: bool = False, stream=None) -> None:
    """EmbeddingForward<InputT,OutputT,IndexT,OffsetT,fp16_math>
    (cuembed/include/embedding_lookup.cuh:245-259).  InputT / OutputT / IndexT /
    OffsetT are taken from the tensor dtypes."""
    lib = _lib.load()
    if weights is not None and weights.dtype != params.dtype:
        raise CuEmbedError("weights must have the element type of params "
                           "(GetElemT<InputT>, embedding_lookup.cuh:249)")
    if _DEBUG_CHECKS:
        DebugCheckLookup(indices, params.shape[0], offsets,
                         batch_size if offsets is not None else None,
                         nnz=None if offsets is not None else batch_size * num_hots,
                         stream=stream)
    _check(lib.cuembed_forward(
        _dev(params, "params"), _dt(params), int(embed_width),
        _dev(indices, "indices"), _it(indices),
        _dev(offsets, "offsets"), _it(offsets) if offsets is not None else 0,
        _dev(weights, "weights"), int(batch_size), int(num_hots), int(mode),
        int(bool(fp16_math)), _dev(ret, "ret"), _dt(ret), _stream(stream)))


def DebugCheckLookup(indices: torch.Tensor, num_rows: int,
                     offsets: Optional[torch.Tensor] = None,
                     batch_size: Optional[int] = None, nnz: Optional[int] = None,
                     stream=None) -> None:
    """cuembed_debug_check_lookup: validates a lookup's indices (all in
    [0, num_rows)) and CSR offsets (non-negative, ascending, within nnz) on the
    device; raises CuEmbedError naming the first offending position.  The
    kernels themselves carry no bounds checks, like the reference's
    (cuembed/include/embedding_lookup_ops.cuh:59).  Synchronises the stream.
    Setting CUEMBED_DEBUG_CHECKS=1 makes EmbeddingForward call it on every
    lookup."""
    lib = _lib.load()
    n = int(indices.numel() if nnz is None else nnz)
    if offsets is not None and batch_size is None:
        batch_size = offsets.numel() - 1
    first = ctypes.c_longlong(-1)
    rc = lib.cuembed_debug_check_lookup(
        _dev(indices, "indices"), _it(indices), n, int(num_rows),
        _dev(offsets, "offsets"), _it(offsets) if offsets is not None else 0,
        int(batch_size or 0), ctypes.byref(first), _stream(stream))
    if rc != 0:
        what = "bag" if rc == -11 else "lookup"
        raise CuEmbedError(f"cuembed_b200 error {rc}: {lib.cuembed_error_string(rc).decode()} "
                           f"(first offending {what}: {first.value})")


def EmbeddingForwardMapped(params: torch.Tensor, embed_width: int,
                           indices: torch.Tensor, offsets: Optional[torch.Tensor],
                           weights: Optional[torch.Tensor], batch_size: int,
                           num_hots: int, mode: CombineMode, ret: torch.Tensor,
                           row_map: torch.Tensor,
                           cache_params: Optional[torch.Tensor] = None,
                           stream=None) -> None:
    """cuembed_forward_mapped: EmbeddingForward through an addresser indirection
    (the reference's embedding-cache hook, embedding_lookup_kernels.cuh:114-115):
    row_map[i] >= 0 reads row row_map[i] of `cache_params` (of `params` when no
    cache table is given), row_map[i] < 0 reads row i of `params`.  `params` only
    has to be device-accessible (e.g. a pinned host tensor mapped into the
    device's address space is the caller's business; here a CUDA tensor)."""
    lib = _lib.load()
    if weights is not None and weights.dtype != params.dtype:
        raise CuEmbedError("weights must have the element type of params")
    if row_map.dtype != indices.dtype:
        raise CuEmbedError("row_map must have the integer type of indices")
    if cache_params is not None and (cache_params.dtype != params.dtype or
                                     cache_params.shape[1] != params.shape[1]):
        raise CuEmbedError("cache_params must have the dtype and row width of params")
    _check(lib.cuembed_forward_mapped(
        _dev(params, "params"), _dt(params), int(embed_width),
        _dev(indices, "indices"), _it(indices),
        _dev(offsets, "offsets"), _it(offsets) if offsets is not None else 0,
        _dev(weights, "weights"), int(batch_size), int(num_hots), int(mode),
        _dev(ret, "ret"), _dt(ret), _dev(row_map, "row_map"),
        _dev(cache_params, "cache_params"), _stream(stream)))


def forward_hot_capacity(dtype: torch.dtype, embed_width: int) -> int:
    """Rows the shared-memory hot-row cache of EmbeddingForwardHot holds for this
    row shape (0: shape not supported)."""
    return int(_lib.load().cuembed_forward_hot_capacity(_DT[dtype], int(embed_width)))


def HotRowsFromSorted(sorted_keys: torch.Tensor, nnz: int, min_count: int,
                      capacity: int, stream=None):
    """Rows hit at least `min_count` times in a sorted (transposed) index array:
    returns (hot_rows int32 [capacity], hot_count int32 [1]) on the device."""
    lib = _lib.load()
    hot_rows = torch.full((max(capacity, 1),), -1, dtype=torch.int32, device=sorted_keys.device)
    hot_count = torch.zeros(1, dtype=torch.int32, device=sorted_keys.device)
    _check(lib.cuembed_hot_rows_from_sorted(
        _dev(sorted_keys, "sorted_keys"), _it(sorted_keys), int(nnz), int(min_count),
        _dev(hot_rows, "hot_rows"), int(capacity), _dev(hot_count, "hot_count"),
        _stream(stream)))
    return hot_rows, hot_count


def EmbeddingForwardHot(params: torch.Tensor, embed_width: int, indices: torch.Tensor,
                        offsets: Optional[torch.Tensor], weights: Optional[torch.Tensor],
                        batch_size: int, num_hots: int, mode: CombineMode,
                        ret: torch.Tensor, hot_rows: torch.Tensor, hot_count: torch.Tensor,
                        stream=None) -> None:
    """cuembed_forward_hot: EmbeddingForward with a shared-memory cache of the
    rows listed in `hot_rows[:hot_count]` (int32, device).  Bit-identical to
    EmbeddingForward for any list; int32 indices, rows of 128 / 256 / 512 bytes."""
    lib = _lib.load()
    if weights is not None and weights.dtype != params.dtype:
        raise CuEmbedError("weights must have the element type of the table")
    if hot_rows.dtype != torch.int32 or hot_count.dtype != torch.int32:
        raise CuEmbedError("hot_rows / hot_count must be int32")
    _check(lib.cuembed_forward_hot(
        _dev(params, "params"), _dt(params), int(embed_width), _dev(indices, "indices"),
        _it(indices), _dev(offsets, "offsets"), _it(offsets) if offsets is not None else 0,
        _dev(weights, "weights"), int(batch_size), int(num_hots), int(mode),
        _dev(ret, "ret"), _dt(ret), _dev(hot_rows, "hot_rows"), _dev(hot_count, "hot_count"),
        int(hot_rows.numel()), _stream(stream)))


def ExtractRowIdsFromFixed(batch_size: int, num_hots: int,
                           row_ids: torch.Tensor, stream=None) -> None:
    """cuembed/include/index_transforms.cuh:45-55."""
    _check(_lib.load().cuembed_extract_row_ids_fixed(
        int(batch_size), int(num_hots), _dev(row_ids, "row_ids"), _it(row_ids),
        _stream(stream)))


def ExtractRowIdsFromCSR(offsets: torch.Tensor, batch_size: int,
                         row_ids: torch.Tensor, stream=None) -> None:
    """cuembed/include/index_transforms.cuh:66-74."""
    _check(_lib.load().cuembed_extract_row_ids_csr(
        _dev(offsets, "offsets"), _it(offsets), int(batch_size),
        _dev(row_ids, "row_ids"), _it(row_ids), _stream(stream)))


def ExtractRowIdsForConcat(nnz: int, row_ids: torch.Tensor, stream=None) -> None:
    """cuembed/include/index_transforms.cuh:85-93."""
    _check(_lib.load().cuembed_extract_row_ids_concat(
        int(nnz), _dev(row_ids, "row_ids"), _it(row_ids), _stream(stream)))


def Transpose(rows: torch.Tensor, cols: torch.Tensor,
              weights: Optional[torch.Tensor], nnz: int,
              transpose_rows: Optional[torch.Tensor],
              transpose_cols: Optional[torch.Tensor],
              transpose_weights: Optional[torch.Tensor],
              work: Optional[torch.Tensor], stream=None) -> int:
    """Transpose<IndexT,WeightT> (cuembed/include/index_transforms.cuh:224-234).
    With work=None this is the workspace query and returns the byte count
    (the reference writes it to *lwork); otherwise `work` is a uint8 CUDA
    tensor of at least that size."""
    lib = _lib.load()
    lwork = ctypes.c_size_t(0 if work is None else work.numel())
    ref = cols if cols is not None else transpose_rows
    wdt = _dt(weights) if weights is not None else 0
    if work is None:
        _check(lib.cuembed_transpose(
            None, None, ctypes.c_void_p(1) if weights is not None else None, wdt,
            int(nnz), _it(ref), None, None, None, None, ctypes.byref(lwork),
            _stream(stream)))
        return int(lwork.value)
    _check(lib.cuembed_transpose(
        _dev(rows, "rows"), _dev(cols, "cols"), _dev(weights, "weights"), wdt,
        int(nnz), _it(cols), _dev(transpose_rows, "transpose_rows"),
        _dev(transpose_cols, "transpose_cols"),
        _dev(transpose_weights, "transpose_weights"), _dev(work, "work"),
        ctypes.byref(lwork), _stream(stream)))
    return int(lwork.value)


def TransposeFixed(cols: torch.Tensor, weights: Optional[torch.Tensor],
                   batch_size: int, num_hots: int,
                   transpose_rows: torch.Tensor, transpose_cols: torch.Tensor,
                   transpose_weights: Optional[torch.Tensor],
                   work: torch.Tensor, stream=None) -> None:
    """cuembed_transpose_fixed: ExtractRowIdsFromFixed + Transpose in one call
    (the sample ids position / num_hots are synthesised in the first sort pass;
    no row-id array).  `work` sized by Transpose(..., work=None) for
    nnz = batch_size * num_hots."""
    lib = _lib.load()
    lwork = ctypes.c_size_t(work.numel())
    wdt = _dt(weights) if weights is not None else 0
    _check(lib.cuembed_transpose_fixed(
        _dev(cols, "cols"), int(batch_size), int(num_hots), _dev(weights, "weights"), wdt,
        _it(cols), _dev(transpose_rows, "transpose_rows"),
        _dev(transpose_cols, "transpose_cols"),
        _dev(transpose_weights, "transpose_weights"), _dev(work, "work"),
        ctypes.byref(lwork), _stream(stream)))


def ComputeCompressedGradIndices(indices: torch.Tensor, nnz: int,
                                 remapped_indices: Optional[torch.Tensor],
                                 work: Optional[torch.Tensor], stream=None) -> int:
    """ComputeCompressedGradIndices<IndexT>
    (cuembed/include/index_transforms.cuh:278-284); work=None is the query."""
    lib = _lib.load()
    lwork = ctypes.c_size_t(0 if work is None else work.numel())
    if work is None:
        _check(lib.cuembed_compressed_grad_indices(
            None, _it(indices), int(nnz), None, None, ctypes.byref(lwork),
            _stream(stream)))
        return int(lwork.value)
    _check(lib.cuembed_compressed_grad_indices(
        _dev(indices, "indices"), _it(indices), int(nnz),
        _dev(remapped_indices, "remapped_indices"), _dev(work, "work"),
        ctypes.byref(lwork), _stream(stream)))
    return int(lwork.value)


def EmbeddingBackward(grad_y: torch.Tensor, embed_width: int,
                      num_grad_embedding_rows: int, nnz: int,
                      transpose_indices: torch.Tensor,
                      transpose_sample_ids: torch.Tensor,
                      transpose_remapped_indices: Optional[torch.Tensor],
                      transpose_weights: Optional[torch.Tensor],
                      skip_grad_init: bool, grad_embedding: torch.Tensor,
                      inverse_mapping: Optional[torch.Tensor],
                      work: Optional[torch.Tensor] = None, stream=None) -> None:
    """EmbeddingBackward<GradT,IndexT>
    (cuembed/include/embedding_lookup.cuh:423-435).  `work` (optional uint8
    CUDA tensor sized by backward_workspace_bytes) selects the explicit
    workspace entry point; without it scratch comes from the library's
    stream-ordered pool."""
    lib = _lib.load()
    if transpose_weights is not None and transpose_weights.dtype != grad_y.dtype:
        raise CuEmbedError("transpose_weights must have the dtype of grad_y")
    if grad_embedding.dtype != grad_y.dtype:
        raise CuEmbedError("grad_embedding must have the dtype of grad_y")
    args = [
        _dev(grad_y, "grad_y"), _dt(grad_y), int(embed_width),
        int(num_grad_embedding_rows), int(nnz), _it(transpose_indices),
        _dev(transpose_indices, "transpose_indices"),
        _dev(transpose_sample_ids, "transpose_sample_ids"),
        _dev(transpose_remapped_indices, "transpose_remapped_indices"),
        _dev(transpose_weights, "transpose_weights"), int(bool(skip_grad_init)),
        _dev(grad_embedding, "grad_embedding"),
        _dev(inverse_mapping, "inverse_mapping"),
    ]
    if work is None:
        _check(lib.cuembed_backward(*args, _stream(stream)))
    else:
        lwork = ctypes.c_size_t(work.numel())
        _check(lib.cuembed_backward_ws(*args, _dev(work, "work"),
                                       ctypes.byref(lwork), _stream(stream)))


def backward_workspace_bytes(dtype: torch.dtype, embed_width: int, nnz: int,
                             index_dtype: torch.dtype = torch.int32) -> int:
    lib = _lib.load()
    lwork = ctypes.c_size_t(0)
    _check(lib.cuembed_backward_ws(
        None, _DT[dtype], int(embed_width), 0, int(nnz), _IT[index_dtype], None,
        None, None, None, 1, None, None, None, ctypes.byref(lwork), None))
    return int(lwork.value)


def EmbeddingForwardMulti(params, embed_width: int, indices, offsets, weights,
                          batch_sizes, num_hots, modes, rets,
                          out_row_stride: int = 0, stream=None) -> None:
    """cuembed_forward_multi: pooled lookups into several tables of one row
    shape in one launch (per 32 tables).  Lists of per-table tensors / ints;
    `offsets` / `weights` may be None (all fixed-hotness / unweighted) or lists
    whose entries may be None (offsets only).  `rets[t]` may be a strided view
    into one [batch, num_tables * embed_width] matrix (pass its row stride in
    elements as out_row_stride).  New functionality: the reference is
    single-table (README.md:110)."""
    lib = _lib.load()
    n = len(params)
    if n == 0:
        return
    vpa = ctypes.c_void_p * n
    ia = ctypes.c_int * n

    def ptrs(ts, name, need_contig=True):
        out = []
        for t in ts:
            if t is None:
                out.append(None)
                continue
            if not t.is_cuda:
                raise CuEmbedError(
                    f"{name} must be CUDA tensors: cuembed_b200 has no CPU fallback")
            if need_contig and not t.is_contiguous():
                raise CuEmbedError(f"{name} must be contiguous")
            out.append(t.data_ptr())
        return vpa(*out)

    for t in rets:
        if t.stride(-1) != 1 or (out_row_stride and t.dim() == 2 and t.shape[0] > 1
                                 and t.stride(0) != out_row_stride):
            raise CuEmbedError("rets must have unit column stride and the given row stride")
    off_list = offsets if offsets is not None else [None] * n
    first_off = next((o for o in off_list if o is not None), None)
    w_arr = ptrs(weights, "weights") if weights is not None else None
    modes_arr = ia(*[int(m) for m in modes]) if modes is not None else None
    _check(lib.cuembed_forward_multi(
        n, ptrs(params, "params"), _dt(params[0]), int(embed_width),
        ptrs(indices, "indices"), _it(indices[0]),
        ptrs(off_list, "offsets") if first_off is not None else None,
        _it(first_off) if first_off is not None else 0, w_arr,
        ia(*[int(b) for b in batch_sizes]), ia(*[int(h) for h in num_hots]),
        modes_arr, ptrs(rets, "rets", need_contig=False), _dt(rets[0]),
        int(out_row_stride), _stream(stream)))


OPT_SGD = 1
OPT_ADAGRAD = 2


def EmbeddingBackwardUpdate(grad_y: torch.Tensor, embed_width: int, nnz: int,
                            transpose_indices: torch.Tensor,
                            transpose_sample_ids: torch.Tensor,
                            transpose_weights: Optional[torch.Tensor],
                            optimizer: int, lr: float, params: torch.Tensor,
                            state: Optional[torch.Tensor] = None,
                            eps: float = 1e-10,
                            work: Optional[torch.Tensor] = None,
                            stream=None) -> None:
    """cuembed_backward_update: backward fused with a sparse optimizer step on
    the touched rows of `params` (SGD, or Adagrad with an fp32 `state` of the
    table's shape).  transpose_indices are table rows (no compressed indices).
    The reference only lists this as a future kernel type (README.md:119)."""
    lib = _lib.load()
    if params.dtype != grad_y.dtype:
        raise CuEmbedError("params must have the dtype of grad_y")
    if transpose_weights is not None and transpose_weights.dtype != grad_y.dtype:
        raise CuEmbedError("transpose_weights must have the dtype of grad_y")
    if optimizer == OPT_ADAGRAD:
        if state is None or state.dtype != torch.float32 or state.shape != params.shape:
            raise CuEmbedError("Adagrad needs an fp32 state of the table's shape")
    head = [_dev(grad_y, "grad_y"), _dt(grad_y), int(embed_width), int(nnz),
            _it(transpose_indices)]
    lwork = ctypes.c_size_t(0)
    _check(lib.cuembed_backward_update(*head, None, None, None, int(optimizer),
                                       float(lr), float(eps), None, None, None,
                                       ctypes.byref(lwork), None))
    if work is None:
        work = torch.empty(lwork.value, dtype=torch.uint8, device=grad_y.device)
    elif work.numel() < lwork.value:
        raise CuEmbedError(f"workspace too small: {work.numel()} < {lwork.value}")
    lwork = ctypes.c_size_t(work.numel())
    _check(lib.cuembed_backward_update(
        *head, _dev(transpose_indices, "transpose_indices"),
        _dev(transpose_sample_ids, "transpose_sample_ids"),
        _dev(transpose_weights, "transpose_weights"), int(optimizer), float(lr),
        float(eps), _dev(params, "params"), _dev(state, "state"),
        _dev(work, "work"), ctypes.byref(lwork), _stream(stream)))


def ShardSelect(indices: torch.Tensor, offsets: Optional[torch.Tensor],
                weights: Optional[torch.Tensor], batch_size: int, num_hots: int,
                row_lo: int, row_hi: int, local_offsets: Optional[torch.Tensor],
                local_indices: Optional[torch.Tensor],
                local_weights: Optional[torch.Tensor],
                work: Optional[torch.Tensor], stream=None) -> int:
    """cuembed_shard_select: compact local CSR of the lookups in
    [row_lo, row_hi).  work=None is the workspace query (returns bytes)."""
    lib = _lib.load()
    lwork = ctypes.c_size_t(0 if work is None else work.numel())
    wdt = _dt(weights) if weights is not None else 0
    if work is None:
        _check(lib.cuembed_shard_select(
            None, _it(indices), ctypes.c_void_p(1) if offsets is not None else None,
            _it(offsets) if offsets is not None else 0, None, wdt, int(batch_size),
            int(num_hots), int(row_lo), int(row_hi), None, None, None, None,
            ctypes.byref(lwork), _stream(stream)))
        return int(lwork.value)
    if local_offsets.dtype != torch.int32:
        raise CuEmbedError("local_offsets must be int32")
    _check(lib.cuembed_shard_select(
        _dev(indices, "indices"), _it(indices), _dev(offsets, "offsets"),
        _it(offsets) if offsets is not None else 0, _dev(weights, "weights"), wdt,
        int(batch_size), int(num_hots), int(row_lo), int(row_hi),
        _dev(local_offsets, "local_offsets"), _dev(local_indices, "local_indices"),
        _dev(local_weights, "local_weights"), _dev(work, "work"),
        ctypes.byref(lwork), _stream(stream)))
    return int(lwork.value)


def ShardSelectCoo(indices: torch.Tensor, offsets: Optional[torch.Tensor],
                   weights: Optional[torch.Tensor], batch_size: int, num_hots: int,
                   row_lo: int, row_hi: int, counts: Optional[torch.Tensor],
                   local_offsets: torch.Tensor, local_indices: torch.Tensor,
                   local_sample_ids: Optional[torch.Tensor],
                   local_weights: Optional[torch.Tensor], work: torch.Tensor,
                   stream=None) -> None:
    """cuembed_shard_select_coo: like ShardSelect, optionally reusing per-bag
    counts and writing the sample id of every selected lookup."""
    lib = _lib.load()
    lwork = ctypes.c_size_t(work.numel())
    if local_offsets.dtype != torch.int32:
        raise CuEmbedError("local_offsets must be int32")
    if counts is not None and counts.dtype != torch.int32:
        raise CuEmbedError("counts must be int32")
    if local_sample_ids is not None and local_sample_ids.dtype != indices.dtype:
        raise CuEmbedError("local_sample_ids must have the dtype of indices")
    _check(lib.cuembed_shard_select_coo(
        _dev(indices, "indices"), _it(indices), _dev(offsets, "offsets"),
        _it(offsets) if offsets is not None else 0, _dev(weights, "weights"),
        _dt(weights) if weights is not None else 0, int(batch_size), int(num_hots),
        int(row_lo), int(row_hi), _dev(counts, "counts"),
        _dev(local_offsets, "local_offsets"), _dev(local_indices, "local_indices"),
        _dev(local_sample_ids, "local_sample_ids"), _dev(local_weights, "local_weights"),
        _dev(work, "work"), ctypes.byref(lwork), _stream(stream)))


def ShardFinalize(partial: torch.Tensor, n_samples: int, embed_width: int,
                  mode: CombineMode, offsets: Optional[torch.Tensor],
                  num_hots: int, sample0: int, weights: Optional[torch.Tensor],
                  out: torch.Tensor, stream=None) -> None:
    """cuembed_shard_finalize: epilogue after the reduce-scatter."""
    if partial.dtype != torch.float32:
        raise CuEmbedError("partial sums must be float32")
    _check(_lib.load().cuembed_shard_finalize(
        _dev(partial, "partial"), int(n_samples), int(embed_width), int(mode),
        _dev(offsets, "offsets"), _it(offsets) if offsets is not None else 0,
        int(num_hots), int(sample0), _dev(weights, "weights"),
        _dt(weights) if weights is not None else 0, _dev(out, "out"), _dt(out),
        _stream(stream)))
