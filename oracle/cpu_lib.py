"""ctypes front-end for the CPU checkers.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Loads either
  * oracle/libcuembed_oracle.so  -- the plain-C restatement (prefix ``oracle_``)
  * oracle/_ref/libcuembed_ref.so -- the reference's own CPU templates compiled
    from /root/reference (prefix ``ref_``), when it has been built.
Both export the same signatures, so ``CpuLib("oracle")`` and ``CpuLib("ref")``
are interchangeable.  Only tests/, bench.py's CPU-baseline legs and
__graft_entry__.smoke() import this module.

Arrays are numpy; float16 is np.float16; bfloat16 is carried as np.uint16 bit
patterns with dtype code 2 passed explicitly (``Bf16`` wrapper below).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

F32, F16, BF16 = 0, 1, 2
SUM, MEAN, CONCAT = 0, 1, 2


class Bf16:
    """A uint16 numpy array whose bits are bfloat16 values."""

    def __init__(self, bits: np.ndarray):
        assert bits.dtype == np.uint16
        self.bits = np.ascontiguousarray(bits)

    @staticmethod
    def from_f32(x: np.ndarray) -> "Bf16":
        u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
        lsb = (u >> 16) & 1
        r = ((u + 0x7FFF + lsb) >> 16).astype(np.uint16)
        return Bf16(r.reshape(x.shape))

    def to_f32(self) -> np.ndarray:
        return (self.bits.astype(np.uint32) << 16).view(np.float32)

    @property
    def shape(self):
        return self.bits.shape


def dt_code(a) -> int:
    if isinstance(a, Bf16):
        return BF16
    if a.dtype == np.float32:
        return F32
    if a.dtype == np.float16:
        return F16
    raise TypeError(f"unsupported element dtype {a.dtype}")


def it_code(a: np.ndarray) -> int:
    if a.dtype == np.int32:
        return 0
    if a.dtype == np.int64:
        return 1
    raise TypeError(f"unsupported index dtype {a.dtype}")


def _raw(a):
    return a.bits if isinstance(a, Bf16) else a


def _ptr(a):
    if a is None:
        return None
    a = _raw(a)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def empty_like_dt(shape, dt: int):
    if dt == F32:
        return np.zeros(shape, np.float32)
    if dt == F16:
        return np.zeros(shape, np.float16)
    return Bf16(np.zeros(shape, np.uint16))


def build(ref: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    target = ["libcuembed_oracle.so"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", _HERE, *target], check=True, capture_output=True)


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libcuembed_ref.so"))


class CpuLib:
    def __init__(self, kind: str = "oracle"):
        assert kind in ("oracle", "ref")
        self.kind = kind
        path = (
            os.path.join(_HERE, "libcuembed_oracle.so")
            if kind == "oracle"
            else os.path.join(_HERE, "_ref", "libcuembed_ref.so")
        )
        if not os.path.exists(path):
            if kind == "oracle":
                build(ref=False)
            else:
                raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        p = kind + "_"
        vp, ci = ctypes.c_void_p, ctypes.c_int
        self._fwd = getattr(self.lib, p + "forward")
        self._fwd.restype = ci
        self._fwd.argtypes = [vp, ci, ci, ci, ci, vp, ci, vp, ci, vp, vp, ci, ci, ci, ci, ci]
        self._xf = getattr(self.lib, p + "extract_row_ids_fixed")
        self._xf.restype = None
        self._xf.argtypes = [ci, ci, vp, ci]
        self._xc = getattr(self.lib, p + "extract_row_ids_csr")
        self._xc.restype = None
        self._xc.argtypes = [vp, ci, ci, vp, ci]
        self._xk = getattr(self.lib, p + "extract_row_ids_concat")
        self._xk.restype = None
        self._xk.argtypes = [ci, vp, ci]
        self._cg = getattr(self.lib, p + "compressed_grad_indices")
        self._cg.restype = None
        self._cg.argtypes = [vp, ci, ci, vp]
        self._tr = getattr(self.lib, p + "transpose")
        self._tr.restype = ci
        self._tr.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
        self._bw = getattr(self.lib, p + "backward")
        self._bw.restype = ci
        self._bw.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp, vp, vp, ci, vp, vp, ci, ci, ci]

    # -- forward ---------------------------------------------------------
    def forward(self, params, indices, offsets, weights, batch_size, num_hots,
                mode, embed_width=None, out_dt: Optional[int] = None,
                fp16_math=False, ret=None, sample_begin=0, sample_end=None):
        in_dt = dt_code(params)
        out_dt = in_dt if out_dt is None else out_dt
        if embed_width is None:
            embed_width = params.shape[-1]
        nnz = int(offsets[batch_size]) if offsets is not None else batch_size * num_hots
        if ret is None:
            shape = (nnz, embed_width) if mode == CONCAT else (batch_size, embed_width)
            ret = empty_like_dt(shape, out_dt)
        rc = self._fwd(_ptr(params), in_dt, embed_width, batch_size, num_hots,
                       _ptr(indices), it_code(indices), _ptr(offsets),
                       it_code(offsets) if offsets is not None else 0,
                       _ptr(weights), _ptr(ret), out_dt, mode, int(fp16_math),
                       sample_begin, batch_size if sample_end is None else sample_end)
        if rc != 0:
            raise ValueError(f"{self.kind}_forward rejected the arguments (code {rc})")
        return ret

    # -- index transforms ------------------------------------------------
    def extract_row_ids_fixed(self, batch_size, num_hots, index_dtype):
        out = np.zeros(batch_size * num_hots, index_dtype)
        self._xf(batch_size, num_hots, _ptr(out), it_code(out))
        return out

    def extract_row_ids_csr(self, offsets, batch_size, index_dtype):
        out = np.zeros(int(offsets[batch_size]), index_dtype)
        self._xc(_ptr(offsets), it_code(offsets), batch_size, _ptr(out), it_code(out))
        return out

    def extract_row_ids_concat(self, nnz, index_dtype):
        out = np.zeros(nnz, index_dtype)
        self._xk(nnz, _ptr(out), it_code(out))
        return out

    def compressed_grad_indices(self, indices):
        out = np.zeros_like(indices)
        self._cg(_ptr(indices), it_code(indices), indices.shape[0], _ptr(out))
        return out

    def transpose(self, rows, cols, weights=None):
        nnz = cols.shape[0]
        t_rows = np.zeros_like(cols)
        t_cols = np.zeros_like(cols)
        t_w = None
        wdt = 0
        if weights is not None:
            wdt = dt_code(weights)
            t_w = empty_like_dt(weights.shape, wdt)
        rc = self._tr(_ptr(rows), _ptr(cols), _ptr(weights), wdt, nnz,
                      it_code(cols), _ptr(t_rows), _ptr(t_cols), _ptr(t_w))
        if rc != 0:
            raise ValueError(f"{self.kind}_transpose failed (code {rc})")
        return t_rows, t_cols, t_w

    # -- backward --------------------------------------------------------
    def backward(self, grad_y, embed_width, num_rows, t_indices, t_sample_ids,
                 t_remapped=None, t_weights=None, skip_grad_init=False,
                 grad_embedding=None, inverse_mapping=None, acc_f32=False,
                 nz_begin=0, nz_end=None):
        dt = dt_code(grad_y)
        nnz = t_indices.shape[0]
        if grad_embedding is None:
            grad_embedding = empty_like_dt((num_rows, embed_width), dt)
        if t_remapped is not None and inverse_mapping is None:
            inverse_mapping = np.zeros(num_rows, t_indices.dtype)
        rc = self._bw(_ptr(grad_y), dt, embed_width, num_rows, nnz,
                      it_code(t_indices), _ptr(t_indices), _ptr(t_sample_ids),
                      _ptr(t_remapped), _ptr(t_weights), int(skip_grad_init),
                      _ptr(grad_embedding), _ptr(inverse_mapping), int(acc_f32),
                      nz_begin, nnz if nz_end is None else nz_end)
        if rc != 0:
            raise ValueError(f"{self.kind}_backward failed (code {rc})")
        return grad_embedding, inverse_mapping
