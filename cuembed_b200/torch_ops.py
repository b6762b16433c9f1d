"""PyTorch custom ops + autograd for the embedding path.

Mirrors the reference's PyTorch example -- the `cuembed_pyt` op library of
examples/pytorch/cuembed_embedding.cu:10-190 and the autograd wrapper
`cuemb_embedding` of examples/pytorch/cuembed_pyt.py:1-77 -- on top of the
sm_100a library: same op names, argument order and meaning, so the reference's
own test script (examples/pytorch/cuembed_test.py) runs against it after
`from cuembed_b200.torch_ops import cuemb_embedding`.

    torch.ops.cuembed_pyt.cuembed_extract_row_ids_from_csr(offsets, nnz)
    torch.ops.cuembed_pyt.cuembed_transpose(rows, cols, weights)
    torch.ops.cuembed_pyt.cuembed_embedding_forward(params, indices, offsets, weights, mode)
    torch.ops.cuembed_pyt.cuembed_embedding_backward(y_grad, num_categories, t_idx, t_sid, t_w)

Beyond the reference example (fp32 / int64 / sum only): fp16 and bf16 tables,
int32 or int64 indices and offsets, mode "mean" (autograd folds 1 / bag length
into the COO weights, because the backward never scales, README.md:117), and
`sparse_grad=True`, which returns the compressed gradient as a sparse COO
tensor (unique rows only) instead of a dense [num_categories, width] tensor.
Fake (meta) registrations make every op traceable by torch.compile.

The ops are CUDA-only: they launch kernels from libcuembed_b200.so on the
current stream; there is no CPU implementation.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import api
from .api import CombineMode

_MODES = {"sum": CombineMode.kSum, "mean": CombineMode.kMean}


def _full_offsets(offsets: torch.Tensor, nnz: int) -> torch.Tensor:
    """The reference passes offsets[:-1] and lets the kernel read one element
    past the slice (examples/pytorch/cuembed_pyt.py:23,
    cuembed/include/index_transforms_kernels.cuh:31-33); rebuild the full
    [batch + 1] array instead of relying on that."""
    tail = torch.full((1,), nnz, dtype=offsets.dtype, device=offsets.device)
    return torch.cat([offsets.contiguous(), tail])


# ------------------------------------------------------------------ the ops
@torch.library.custom_op("cuembed_pyt::cuembed_extract_row_ids_from_csr",
                         mutates_args=(), device_types="cuda")
def cuembed_extract_row_ids_from_csr(offsets: torch.Tensor, nnz: int) -> torch.Tensor:
    batch = offsets.size(0)
    row_ids = torch.empty(nnz, dtype=offsets.dtype, device=offsets.device)
    if nnz > 0:
        api.ExtractRowIdsFromCSR(_full_offsets(offsets, nnz), batch, row_ids)
    return row_ids


@cuembed_extract_row_ids_from_csr.register_fake
def _(offsets, nnz):
    return torch.empty((nnz,), device=offsets.device, dtype=offsets.dtype)


@torch.library.custom_op("cuembed_pyt::cuembed_transpose", mutates_args=(),
                         device_types="cuda")
def cuembed_transpose(rows: torch.Tensor, cols: torch.Tensor,
                      weights: Optional[torch.Tensor] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    nnz = rows.size(0)
    rows_c, cols_c = rows.contiguous(), cols.contiguous()
    w_c = weights.contiguous() if weights is not None else None
    t_rows = torch.empty_like(rows_c)
    t_cols = torch.empty_like(cols_c)
    # like the reference, "no weights" comes back as an empty tensor
    t_w = torch.empty(nnz if w_c is not None else 0,
                      dtype=w_c.dtype if w_c is not None else torch.float32,
                      device=rows.device)
    if nnz > 0:
        lwork = api.Transpose(rows_c, cols_c, w_c, nnz, None, None, None, None)
        work = torch.empty(lwork, dtype=torch.uint8, device=rows.device)
        api.Transpose(rows_c, cols_c, w_c, nnz, t_rows, t_cols,
                      t_w if w_c is not None else None, work)
    return t_rows, t_cols, t_w


@cuembed_transpose.register_fake
def _(rows, cols, weights=None):
    n = cols.shape[0] if weights is not None else 0
    dt = weights.dtype if weights is not None else torch.float32
    return (torch.empty_like(rows), torch.empty_like(cols),
            torch.empty((n,), device=rows.device, dtype=dt))


@torch.library.custom_op("cuembed_pyt::cuembed_embedding_forward", mutates_args=(),
                         device_types="cuda")
def cuembed_embedding_forward(params: torch.Tensor, indices: torch.Tensor,
                              offsets: torch.Tensor,
                              weights: Optional[torch.Tensor] = None,
                              mode: str = "sum") -> torch.Tensor:
    if mode not in _MODES:
        raise ValueError("mode must be 'sum' or 'mean'")
    batch = offsets.numel() - 1
    width = params.size(1)
    params_c = params.contiguous()
    out = torch.empty(batch, width, dtype=params.dtype, device=params.device)
    if batch > 0:
        api.EmbeddingForward(params_c, width, indices.contiguous(), offsets.contiguous(),
                             weights.contiguous() if weights is not None else None,
                             batch, 0, _MODES[mode], out)
    return out


@cuembed_embedding_forward.register_fake
def _(params, indices, offsets, weights=None, mode="sum"):
    return torch.empty((offsets.shape[0] - 1, params.shape[1]), device=params.device,
                       dtype=params.dtype)


@torch.library.custom_op("cuembed_pyt::cuembed_embedding_backward", mutates_args=(),
                         device_types="cuda")
def cuembed_embedding_backward(y_grad: torch.Tensor, num_categories: int,
                               transpose_indices: torch.Tensor,
                               transpose_sample_ids: torch.Tensor,
                               transpose_weights: Optional[torch.Tensor] = None
                               ) -> torch.Tensor:
    width = y_grad.size(1)
    nnz = transpose_indices.size(0)
    grad = torch.zeros(num_categories, width, dtype=y_grad.dtype, device=y_grad.device)
    if nnz > 0:
        w = transpose_weights.contiguous() if transpose_weights is not None else None
        api.EmbeddingBackward(y_grad.contiguous(), width, num_categories, nnz,
                              transpose_indices.contiguous(),
                              transpose_sample_ids.contiguous(), None, w,
                              True, grad, None)
    return grad


@cuembed_embedding_backward.register_fake
def _(y_grad, num_categories, transpose_indices, transpose_sample_ids,
      transpose_weights=None):
    return torch.empty((num_categories, y_grad.shape[1]), device=y_grad.device,
                       dtype=y_grad.dtype)


@torch.library.custom_op("cuembed_pyt::cuembed_embedding_backward_compressed",
                         mutates_args=(), device_types="cuda")
def cuembed_embedding_backward_compressed(
        y_grad: torch.Tensor, transpose_indices: torch.Tensor,
        transpose_sample_ids: torch.Tensor,
        transpose_weights: Optional[torch.Tensor] = None
        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Compressed gradient: (grad_rows [num_unique, width], unique table rows
    [num_unique]); ComputeCompressedGradIndices + EmbeddingBackward with
    inverse_mapping (cuembed/README.md:79-87)."""
    width = y_grad.size(1)
    nnz = transpose_indices.size(0)
    idt = transpose_indices.dtype
    if nnz == 0:
        return (torch.zeros(0, width, dtype=y_grad.dtype, device=y_grad.device),
                torch.zeros(0, dtype=idt, device=y_grad.device))
    t_idx = transpose_indices.contiguous()
    lwork = api.ComputeCompressedGradIndices(t_idx, nnz, None, None)
    work = torch.empty(lwork, dtype=torch.uint8, device=y_grad.device)
    remapped = torch.empty_like(t_idx)
    api.ComputeCompressedGradIndices(t_idx, nnz, remapped, work)
    num_unique = int(remapped[-1].item()) + 1  # the caller sizes the gradient
    grad = torch.empty(num_unique, width, dtype=y_grad.dtype, device=y_grad.device)
    inv = torch.empty(num_unique, dtype=idt, device=y_grad.device)
    w = transpose_weights.contiguous() if transpose_weights is not None else None
    api.EmbeddingBackward(y_grad.contiguous(), width, num_unique, nnz, t_idx,
                          transpose_sample_ids.contiguous(), remapped, w, True, grad, inv)
    return grad, inv


@cuembed_embedding_backward_compressed.register_fake
def _(y_grad, transpose_indices, transpose_sample_ids, transpose_weights=None):
    n = torch.library.get_ctx().new_dynamic_size()
    return (torch.empty((n, y_grad.shape[1]), device=y_grad.device, dtype=y_grad.dtype),
            torch.empty((n,), device=y_grad.device, dtype=transpose_indices.dtype))


def _backward_update(y_grad, params, state, t_idx, t_sid, t_w, lr, eps, opt):
    if not params.is_contiguous():
        raise ValueError("params must be contiguous (it is updated in place)")
    nnz = t_idx.size(0)
    if nnz == 0:
        return
    w = t_w.contiguous() if t_w is not None else None
    api.EmbeddingBackwardUpdate(y_grad.contiguous(), params.size(1), nnz, t_idx.contiguous(),
                                t_sid.contiguous(), w, opt, lr, params, state=state, eps=eps)


# Backward fused with the sparse optimizer step (an addition,
# the reference lists it as future work, README.md:119): the gradient of the
# touched rows is applied to `params` in place, no gradient tensor exists.
@torch.library.custom_op("cuembed_pyt::cuembed_embedding_backward_sgd",
                         mutates_args=("params",), device_types="cuda")
def cuembed_embedding_backward_sgd(
        y_grad: torch.Tensor, params: torch.Tensor, transpose_indices: torch.Tensor,
        transpose_sample_ids: torch.Tensor, transpose_weights: Optional[torch.Tensor],
        lr: float) -> None:
    _backward_update(y_grad, params, None, transpose_indices, transpose_sample_ids,
                     transpose_weights, lr, 0.0, api.OPT_SGD)


@cuembed_embedding_backward_sgd.register_fake
def _(y_grad, params, transpose_indices, transpose_sample_ids, transpose_weights, lr):
    return None


@torch.library.custom_op("cuembed_pyt::cuembed_embedding_backward_adagrad",
                         mutates_args=("params", "state"), device_types="cuda")
def cuembed_embedding_backward_adagrad(
        y_grad: torch.Tensor, params: torch.Tensor, state: torch.Tensor,
        transpose_indices: torch.Tensor, transpose_sample_ids: torch.Tensor,
        transpose_weights: Optional[torch.Tensor], lr: float, eps: float) -> None:
    _backward_update(y_grad, params, state, transpose_indices, transpose_sample_ids,
                     transpose_weights, lr, eps, api.OPT_ADAGRAD)


@cuembed_embedding_backward_adagrad.register_fake
def _(y_grad, params, state, transpose_indices, transpose_sample_ids,
      transpose_weights, lr, eps):
    return None


def cuemb_embedding_sgd_step(params, idx, offsets, out_grad, lr: float, weights=None,
                             mode: str = "sum", optimizer: str = "sgd", state=None,
                             eps: float = 1e-10) -> None:
    """One sparse optimizer step of an EmbeddingBag(include_last_offset=True)
    table given d loss / d output: row ids -> transpose -> fused update, without
    materialising the gradient (use instead of autograd + torch.optim for the
    embedding table).  `params` is updated in place."""
    nnz = idx.size(0)
    sample_ids = cuembed_extract_row_ids_from_csr(offsets[:-1], nnz)
    if sample_ids.dtype != idx.dtype:
        sample_ids = sample_ids.to(idx.dtype)
    if mode == "mean" and weights is not None:
        raise ValueError("weighted mean has no backward formula (README.md:117)")
    w = _coo_weights(mode, offsets, weights, sample_ids, out_grad.dtype)
    t_idx, t_sid, t_w = cuembed_transpose(sample_ids, idx, w)
    t_w = None if t_w.numel() == 0 else t_w
    if optimizer == "sgd":
        cuembed_embedding_backward_sgd(out_grad, params, t_idx, t_sid, t_w, lr)
    elif optimizer == "adagrad":
        if state is None:
            raise ValueError("adagrad needs an fp32 state tensor shaped like params")
        cuembed_embedding_backward_adagrad(out_grad, params, state, t_idx, t_sid, t_w,
                                           lr, eps)
    else:
        raise ValueError("optimizer must be 'sgd' or 'adagrad'")


# ----------------------------------------------------------------- autograd
def cuembed_forward(params, idx, offsets, weights, mode="sum"):
    return cuembed_embedding_forward(params, idx, offsets, weights, mode)


def _coo_weights(ctx_mode, offsets, weights, sample_ids, dtype):
    """Per-lookup factor of d out[sample] / d row: the weight (sum) or
    1 / bag length (mean) -- the backward kernel itself never scales."""
    if ctx_mode == "sum":
        return weights
    lens = (offsets[1:] - offsets[:-1]).to(torch.float32)
    inv = torch.where(lens > 0, 1.0 / lens, torch.zeros_like(lens)).to(dtype)
    return inv[sample_ids]


class _CuEmbEmbedding(torch.autograd.Function):
    """examples/pytorch/cuembed_pyt.py:37-51."""

    @staticmethod
    def forward(ctx, params, idx, offsets, weights, mode, sparse_grad):
        ctx.save_for_backward(idx, offsets, weights)
        ctx.num_categories = params.size(0)
        ctx.mode = mode
        ctx.sparse_grad = sparse_grad
        return cuembed_forward(params, idx, offsets, weights, mode)

    @staticmethod
    def backward(ctx, out_grad):
        idx, offsets, weights = ctx.saved_tensors
        nnz = idx.size(0)
        # include_last_offset=True convention, as in the reference
        sample_ids = cuembed_extract_row_ids_from_csr(offsets[:-1], nnz)
        if sample_ids.dtype != idx.dtype:
            sample_ids = sample_ids.to(idx.dtype)
        w = _coo_weights(ctx.mode, offsets, weights, sample_ids, out_grad.dtype)
        t_idx, t_sid, t_w = cuembed_transpose(sample_ids, idx, w)
        t_w = None if t_w.numel() == 0 else t_w
        if ctx.sparse_grad:
            rows, unique = cuembed_embedding_backward_compressed(out_grad, t_idx, t_sid, t_w)
            grad = torch.sparse_coo_tensor(unique.view(1, -1).to(torch.int64), rows,
                                           (ctx.num_categories, out_grad.size(1)),
                                           is_coalesced=True)
        else:
            grad = cuembed_embedding_backward(out_grad, ctx.num_categories, t_idx, t_sid, t_w)
        return grad, None, None, None, None, None


def cuemb_embedding(params, idx, offsets, weights=None, mode: str = "sum",
                    sparse_grad: bool = False):
    """Drop-in for the reference's cuemb_embedding (EmbeddingBag with
    include_last_offset=True): pooled lookup with autograd."""
    if mode == "mean" and weights is not None:
        raise ValueError("weighted mean has no autograd formula here "
                         "(the reference's backward supports sum only, README.md:117)")
    if not torch.is_grad_enabled() or not params.requires_grad:
        return cuembed_forward(params, idx, offsets, weights, mode)
    return _CuEmbEmbedding.apply(params, idx, offsets, weights, mode, sparse_grad)
