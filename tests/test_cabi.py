"""The C-ABI library loads and exports every symbol include/cuembed_b200.h
declares (no compute calls: runs without a GPU); argument checks that do not
touch the device return the documented codes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "cuembed_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cuembed_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for want in ["cuembed_forward", "cuembed_extract_row_ids_fixed",
                 "cuembed_extract_row_ids_csr", "cuembed_extract_row_ids_concat",
                 "cuembed_transpose", "cuembed_compressed_grad_indices",
                 "cuembed_backward", "cuembed_backward_ws"]:
        assert want in names


def test_library_exports_every_declared_symbol():
    from cuembed_b200 import _lib
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # and the Python binding table covers the same set
    assert sorted(_lib.SIGNATURES) == declared_functions()
    assert lib.cuembed_build_arch() == b"sm_100a"
    assert lib.cuembed_version() >= 100


def test_argument_checks_without_device_work():
    """Checks that fail before any CUDA call (reference: CUEMBED_ASSERT,
    cuembed/include/embedding_lookup.cuh:260-267,163)."""
    from cuembed_b200 import _lib
    lib = _lib.load()
    p = ctypes.c_void_p(256)  # never dereferenced: the checks come first
    # weights with concat
    assert lib.cuembed_forward(p, 0, 4, p, 0, None, 0, p, 2, 2, 2, 0, p, 0, None) == -1
    # neither / both of CSR and fixed hotness
    assert lib.cuembed_forward(p, 0, 4, p, 0, p, 0, None, 2, 2, 0, 0, p, 0, None) == -2
    assert lib.cuembed_forward(p, 0, 4, p, 0, None, 0, None, 2, 0, 0, 0, p, 0, None) == -2
    # CSR concat
    assert lib.cuembed_forward(p, 0, 4, p, 0, p, 0, None, 2, 0, 2, 0, p, 0, None) == -3
    # fp16 row of 3 elements = 6 bytes
    assert lib.cuembed_forward(p, 1, 3, p, 0, None, 0, None, 2, 2, 0, 0, p, 1, None) == -4
    # workspace queries work without a device
    lw = ctypes.c_size_t(0)
    assert lib.cuembed_transpose(None, None, None, 0, 1 << 20, 0, None, None, None,
                                 None, ctypes.byref(lw), None) == 0
    assert lw.value >= 2 * 4 * (1 << 20)
    small = ctypes.c_size_t(16)
    assert lib.cuembed_transpose(p, p, None, 0, 1 << 20, 0, p, p, None, p,
                                 ctypes.byref(small), None) == -6
    assert lib.cuembed_transpose(None, None, None, 0, 1 << 30, 0, None, None, None,
                                 None, ctypes.byref(lw), None) == -9
    assert lib.cuembed_compressed_grad_indices(None, 1, 12345, None, None,
                                               ctypes.byref(lw), None) == 0
    assert lw.value > 0
    assert lib.cuembed_backward_ws(None, 1, 256, 0, 4194304, 0, None, None, None, None,
                                   1, None, None, None, ctypes.byref(lw), None) == 0
    assert lw.value > 0
    assert b"Check failed" in lib.cuembed_error_string(-2)


def test_python_api_refuses_cpu_tensors():
    import torch
    import cuembed_b200 as ce
    t = torch.zeros(4, 4)
    idx = torch.zeros(4, dtype=torch.int32)
    with pytest.raises(ce.CuEmbedError, match="no CPU fallback"):
        ce.EmbeddingForward(t, 4, idx, None, None, 2, 2, ce.CombineMode.kSum, torch.zeros(2, 4))
