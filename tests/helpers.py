"""Shared helpers for the parity tests: problem construction (numpy), the CPU
checker pipeline, dtype plumbing.  Test infrastructure only."""
from __future__ import annotations

import numpy as np

from cuembed_b200 import datagen
from oracle import cpu_lib
from oracle.cpu_lib import BF16, CONCAT, F16, F32, MEAN, SUM, Bf16

MODES = {"sum": SUM, "mean": MEAN, "concat": CONCAT}


def cast_elems(x_f32: np.ndarray, dt: int):
    """float32 host array -> array of dtype code dt (Bf16 wrapper for bf16)."""
    if dt == F32:
        return np.ascontiguousarray(x_f32, dtype=np.float32)
    if dt == F16:
        return np.ascontiguousarray(x_f32.astype(np.float16))
    return Bf16.from_f32(np.ascontiguousarray(x_f32, dtype=np.float32))


def raw(a):
    return a.bits if isinstance(a, Bf16) else a


def to_f32(a) -> np.ndarray:
    if isinstance(a, Bf16):
        return a.to_f32()
    return np.asarray(a, dtype=np.float32)


def bits_equal(a, b) -> bool:
    ra, rb = raw(a), raw(b)
    if ra.dtype == np.float32:
        return np.array_equal(ra.view(np.uint32), rb.view(np.uint32))
    if ra.dtype == np.float16:
        return np.array_equal(ra.view(np.uint16), rb.view(np.uint16))
    return np.array_equal(ra, rb)


def value_equal(a, b) -> bool:
    """Equality in the EXPECT_EQ sense (-0 == +0)."""
    return np.array_equal(to_f32(a), to_f32(b))


class Problem:
    """A forward/transpose/backward problem in host memory."""

    def __init__(self, batch, width, hot, mode="sum", csr=False, weighted=False,
                 compressed=False, num_categories=20 * 1024, dt=F32,
                 index_dtype=np.int32, offset_dtype=np.int32, alpha=0.0,
                 seed=7, integer_table=False):
        self.mode_name = mode
        self.mode = MODES[mode]
        self.csr, self.weighted, self.compressed = csr, weighted, compressed
        self.batch, self.width, self.hot = batch, width, hot
        self.num_categories = num_categories
        self.dt = dt
        wl = datagen.make_workload(num_categories, width, batch, hot, alpha=alpha,
                                   csr=csr, weighted=weighted,
                                   index_dtype=index_dtype,
                                   offset_dtype=offset_dtype, seed=seed)
        self.indices = wl.indices
        self.offsets = wl.offsets
        self.nnz = wl.nnz
        self.num_hots = 0 if csr else hot
        table = datagen.make_table(num_categories, width, seed=seed + 1)
        if integer_table:
            table = np.round(table * 8.0)
        self.table = cast_elems(table, dt)
        self.weights = cast_elems(wl.weights, dt) if weighted else None
        gy_rows = self.nnz if mode == "concat" else batch
        self.grad_y = cast_elems(datagen.make_grad_y(gy_rows, width, seed=seed + 2), dt)

    # CPU checker pipeline ------------------------------------------------
    def cpu_forward(self, lib, fp16_math=False, out_dt=None):
        return lib.forward(self.table, self.indices, self.offsets, self.weights,
                           self.batch, self.num_hots, self.mode,
                           embed_width=self.width, fp16_math=fp16_math,
                           out_dt=out_dt)

    def cpu_row_ids(self, lib):
        if self.mode == CONCAT:
            return lib.extract_row_ids_concat(self.nnz, self.indices.dtype)
        if self.csr:
            return lib.extract_row_ids_csr(self.offsets, self.batch, self.indices.dtype)
        return lib.extract_row_ids_fixed(self.batch, self.hot, self.indices.dtype)

    def cpu_transpose(self, lib):
        rows = self.cpu_row_ids(lib)
        t_idx, t_sid, t_w = lib.transpose(rows, self.indices, self.weights)
        remapped = lib.compressed_grad_indices(t_idx) if self.compressed else None
        return rows, t_idx, t_sid, t_w, remapped

    def cpu_backward(self, lib, t_idx, t_sid, t_w, remapped, acc_f32=False,
                     skip_grad_init=False, grad_embedding=None):
        if self.compressed:
            num_rows = int(remapped[-1]) + 1 if self.nnz > 0 else 0
        else:
            num_rows = self.num_categories
        return lib.backward(self.grad_y, self.width, num_rows, t_idx, t_sid,
                            remapped, t_w, skip_grad_init=skip_grad_init,
                            grad_embedding=grad_embedding, acc_f32=acc_f32), num_rows
