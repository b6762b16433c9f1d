"""GPU-side helpers for the parity tests: numpy <-> torch CUDA plumbing and the
CUDA pipeline (through the C ABI via cuembed_b200.api)."""
from __future__ import annotations

import numpy as np
import torch

import cuembed_b200 as ce
from helpers import raw
from oracle.cpu_lib import BF16, CONCAT, F16, F32, Bf16

DEV = "cuda:0"
TORCH_DT = {F32: torch.float32, F16: torch.float16, BF16: torch.bfloat16}


def to_dev(a):
    """numpy / Bf16 host array -> CUDA tensor with the matching dtype."""
    if a is None:
        return None
    if isinstance(a, Bf16):
        t = torch.from_numpy(a.bits.view(np.int16).copy()).to(DEV)
        return t.view(torch.bfloat16)
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def to_host(t: torch.Tensor):
    """CUDA tensor -> numpy (Bf16 wrapper for bfloat16)."""
    if t.dtype == torch.bfloat16:
        return Bf16(t.view(torch.int16).cpu().numpy().view(np.uint16).copy())
    return t.cpu().numpy()


def gpu_forward(p, fp16_math=False, out_dt=None):
    out_dt = p.dt if out_dt is None else out_dt
    rows = p.nnz if p.mode == CONCAT else p.batch
    # Poison the output so that unwritten elements are caught.
    ret = torch.full((rows, p.width), float("nan"), dtype=TORCH_DT[out_dt], device=DEV)
    ce.EmbeddingForward(to_dev(p.table), p.width, to_dev(p.indices),
                        to_dev(p.offsets), to_dev(p.weights), p.batch,
                        p.num_hots, ce.CombineMode(p.mode), ret,
                        fp16_math=fp16_math)
    torch.cuda.synchronize()
    return to_host(ret)


def gpu_row_ids(p):
    idx_t = torch.int32 if p.indices.dtype == np.int32 else torch.int64
    row_ids = torch.full((p.nnz,), -1, dtype=idx_t, device=DEV)
    if p.mode == CONCAT:
        ce.ExtractRowIdsForConcat(p.nnz, row_ids)
    elif p.csr:
        ce.ExtractRowIdsFromCSR(to_dev(p.offsets), p.batch, row_ids)
    else:
        ce.ExtractRowIdsFromFixed(p.batch, p.hot, row_ids)
    return row_ids


def gpu_transpose(p):
    """Returns device tensors (rows, t_idx, t_sid, t_w, remapped)."""
    rows = gpu_row_ids(p)
    idx = to_dev(p.indices)
    w = to_dev(p.weights)
    nnz = p.nnz
    t_idx = torch.full_like(idx, -1)
    t_sid = torch.full_like(idx, -1)
    t_w = torch.zeros_like(w) if w is not None else None
    lwork = ce.Transpose(rows, idx, w, nnz, None, None, None, None)
    lwork = max(lwork, ce.ComputeCompressedGradIndices(idx, nnz, None, None))
    work = torch.empty(lwork, dtype=torch.uint8, device=DEV)
    ce.Transpose(rows, idx, w, nnz, t_idx, t_sid, t_w, work)
    remapped = None
    if p.compressed:
        remapped = torch.full_like(idx, -1)
        ce.ComputeCompressedGradIndices(t_idx, nnz, remapped, work)
    torch.cuda.synchronize()
    return rows, t_idx, t_sid, t_w, remapped


def gpu_backward(p, t_idx, t_sid, t_w, remapped, skip_grad_init=False,
                 explicit_workspace=False, prefill=None):
    nnz = p.nnz
    if p.compressed:
        num_rows = int(remapped[-1].item()) + 1 if nnz > 0 else 0
    else:
        num_rows = p.num_categories
    dt = TORCH_DT[p.dt]
    if prefill is None:
        prefill = 0.0 if skip_grad_init else float("nan")
    grad = torch.full((num_rows, p.width), prefill, dtype=dt, device=DEV)
    inv = None
    if p.compressed:
        inv = torch.full((num_rows,), -1, dtype=t_idx.dtype, device=DEV)
    work = None
    if explicit_workspace:
        nbytes = ce.backward_workspace_bytes(dt, p.width, nnz, t_idx.dtype)
        work = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    ce.EmbeddingBackward(to_dev(p.grad_y), p.width, num_rows, nnz, t_idx, t_sid,
                         remapped, t_w, skip_grad_init, grad, inv, work=work)
    torch.cuda.synchronize()
    return to_host(grad), (inv.cpu().numpy() if inv is not None else None), num_rows
