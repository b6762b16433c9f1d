"""ctypes binding of libcuembed_b200.so (the C ABI in include/cuembed_b200.h).

The library is the product: if it is missing it is built once with nvcc, and
if that fails the import raises -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

_LIB = None

# Names and signatures exactly as declared in include/cuembed_b200.h.
_vp, _ci, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
_szp = ctypes.POINTER(ctypes.c_size_t)
SIGNATURES = {
    "cuembed_version": (_ci, []),
    "cuembed_build_arch": (ctypes.c_char_p, []),
    "cuembed_error_string": (ctypes.c_char_p, [_ci]),
    "cuembed_forward": (_ci, [_vp, _ci, _ci, _vp, _ci, _vp, _ci, _vp, _ci, _ci,
                              _ci, _ci, _vp, _ci, _vp]),
    "cuembed_forward_mapped": (_ci, [_vp, _ci, _ci, _vp, _ci, _vp, _ci, _vp, _ci, _ci,
                                     _ci, _vp, _ci, _vp, _vp, _vp]),
    "cuembed_forward_multi": (_ci, [_ci, ctypes.POINTER(_vp), _ci, _ci,
                                    ctypes.POINTER(_vp), _ci, ctypes.POINTER(_vp), _ci,
                                    ctypes.POINTER(_vp), ctypes.POINTER(_ci),
                                    ctypes.POINTER(_ci), ctypes.POINTER(_ci),
                                    ctypes.POINTER(_vp), _ci, ctypes.c_longlong, _vp]),
    "cuembed_forward_hot_capacity": (_ci, [_ci, _ci]),
    "cuembed_forward_hot": (_ci, [_vp, _ci, _ci, _vp, _ci, _vp, _ci, _vp, _ci, _ci, _ci,
                                  _vp, _ci, _vp, _vp, _ci, _vp]),
    "cuembed_hot_rows_from_sorted": (_ci, [_vp, _ci, _ci, _ci, _vp, _ci, _vp, _vp]),
    "cuembed_extract_row_ids_fixed": (_ci, [_ci, _ci, _vp, _ci, _vp]),
    "cuembed_extract_row_ids_csr": (_ci, [_vp, _ci, _ci, _vp, _ci, _vp]),
    "cuembed_extract_row_ids_concat": (_ci, [_ci, _vp, _ci, _vp]),
    "cuembed_transpose": (_ci, [_vp, _vp, _vp, _ci, _ci, _ci, _vp, _vp, _vp,
                                _vp, _szp, _vp]),
    "cuembed_transpose_fixed": (_ci, [_vp, _ci, _ci, _vp, _ci, _ci, _vp, _vp, _vp, _vp, _szp, _vp]),
    "cuembed_compressed_grad_indices": (_ci, [_vp, _ci, _ci, _vp, _vp, _szp, _vp]),
    "cuembed_backward": (_ci, [_vp, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _vp, _vp,
                               _ci, _vp, _vp, _vp]),
    "cuembed_backward_ws": (_ci, [_vp, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _vp,
                                  _vp, _ci, _vp, _vp, _vp, _szp, _vp]),
    "cuembed_backward_update": (_ci, [_vp, _ci, _ci, _ci, _ci, _vp, _vp, _vp, _ci,
                                      ctypes.c_float, ctypes.c_float, _vp, _vp, _vp,
                                      _szp, _vp]),
    "cuembed_shard_select": (_ci, [_vp, _ci, _vp, _ci, _vp, _ci, _ci, _ci,
                                   ctypes.c_longlong, ctypes.c_longlong, _vp, _vp,
                                   _vp, _vp, _szp, _vp]),
    "cuembed_shard_select_coo": (_ci, [_vp, _ci, _vp, _ci, _vp, _ci, _ci, _ci,
                                       ctypes.c_longlong, ctypes.c_longlong, _vp, _vp,
                                       _vp, _vp, _vp, _vp, _szp, _vp]),
    "cuembed_shard_finalize": (_ci, [_vp, _ci, _ci, _ci, _vp, _ci, _ci, _ci, _vp,
                                     _ci, _vp, _ci, _vp]),
    "cuembed_peer_alloc": (_ci, [_sz, ctypes.POINTER(_vp)]),
    "cuembed_peer_free": (_ci, [_vp]),
    "cuembed_peer_export": (_ci, [_vp, ctypes.c_char_p]),
    "cuembed_peer_open": (_ci, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "cuembed_peer_close": (_ci, [_vp]),
    "cuembed_shard_pool_push": (_ci, [_vp, _ci, _ci, _vp, _ci, _vp, _ci, _vp, _ci, _ci,
                                      ctypes.c_longlong, ctypes.c_longlong,
                                      ctypes.POINTER(_vp), _ci, _ci, _ci, _vp, _vp]),
    "cuembed_shard_concat_push": (_ci, [_vp, _ci, _ci, _vp, _ci, _ci, _ci,
                                        ctypes.c_longlong, ctypes.c_longlong,
                                        ctypes.POINTER(_vp), _ci, _ci, _vp]),
    "cuembed_shard_signal": (_ci, [ctypes.POINTER(_vp), _ci, _ci, _ci, ctypes.c_uint, _vp]),
    "cuembed_shard_wait": (_ci, [_vp, _ci, _ci, ctypes.c_uint, _vp, _sz, _vp]),
    "cuembed_shard_set_timeout_ms": (_ci, [ctypes.c_longlong]),
    "cuembed_shard_reduce_finalize": (_ci, [_vp, _ci, _ci, _vp, _ci, ctypes.c_uint, _ci,
                                            _ci, _ci, _vp, _ci, _ci, _ci, _vp, _ci, _vp,
                                            _ci, _vp]),
    "cuembed_shard_allgather_push": (_ci, [_vp, _sz, ctypes.POINTER(_vp), _ci, _ci, _vp]),
    "cuembed_microbench_gather": (_ci, [_vp, _ci, _vp, ctypes.c_longlong, _ci, _vp, _vp]),
    "cuembed_microbench_gather_bulk": (_ci, [_vp, _ci, _vp, ctypes.c_longlong, _ci, _vp, _vp]),
    "cuembed_debug_check_lookup": (_ci, [_vp, _ci, ctypes.c_longlong, ctypes.c_longlong, _vp, _ci, _ci,
                                         ctypes.POINTER(ctypes.c_longlong), _vp]),
    "cuembed_microbench_gather_async": (_ci, [_vp, _ci, _vp, ctypes.c_longlong, _ci, _vp, _vp]),
    "cuembed_launch_count": (ctypes.c_ulonglong, []),
}


def lib_path() -> str:
    # CUEMBED_B200_LIB: a differently tuned build of the same sources (kernel
    # tuning sweeps under gpurun); the default is the in-tree product library
    return os.environ.get("CUEMBED_B200_LIB") or _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building if necessary) the CUDA library.  Raises on failure."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        _build.build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
