import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_lib
    cpu_lib.build(ref=os.path.isdir("/root/reference/utils/include"))
    return cpu_lib.CpuLib("oracle")


@pytest.fixture(scope="session")
def reflib():
    """The reference's own CPU templates (oracle/_ref), if built."""
    from oracle import cpu_lib
    cpu_lib.build(ref=os.path.isdir("/root/reference/utils/include"))
    if not cpu_lib.ref_available():
        pytest.skip("oracle/_ref/libcuembed_ref.so not built (reference tree absent)")
    return cpu_lib.CpuLib("ref")


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail loudly if it cannot be loaded."""
    import torch
    assert torch.cuda.is_available(), "GPU test without a CUDA device"
    from cuembed_b200 import _lib
    return _lib.load()
