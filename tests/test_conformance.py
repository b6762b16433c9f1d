"""The reference's OWN result-pinning gtest suites, compiled unchanged against
the drop-in (oracle/conformance.sh, top-level CMakeLists.txt with
-DCUEMBED_CONFORMANCE_REFERENCE_DIR) and run on the GPU:

    /root/reference/tests/test_embedding_forward.cu      (KATs, :120-160)
    /root/reference/tests/test_embedding_transpose.cu    (KATs, :112-122)
    /root/reference/tests/test_embedding_backward.cu     (KATs, :162-202)
    /root/reference/tests/test_embedding_against_cpu.cu  (57 shapes x 8 type sets, :236-293)

Their sources, the harness they link (utils/src/*) and gtest / abseil are the
reference's; only "cuembed/include/*" resolves to this repository.  The
binaries are built where /root/reference exists (this container) into
oracle/_ref/conformance/ and travel to the GPU box with the snapshot.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONF = os.path.join(ROOT, "oracle", "_ref", "conformance")
SUITES = ["test_embedding_forward", "test_embedding_transpose",
          "test_embedding_backward", "test_embedding_against_cpu"]
BINARIES = SUITES + ["manual_benchmark"]


def _need(name):
    path = os.path.join(CONF, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (oracle/conformance.sh needs /root/reference)")
    return path


@pytest.mark.parametrize("name", BINARIES)
def test_conformance_binary_binds_the_c_abi(name):
    """The unchanged reference sources resolved to THIS library: they import
    the C ABI and contain none of the reference's kernels / CUB sort calls."""
    path = _need(name)
    syms = subprocess.run(["nm", "-C", path], capture_output=True, text=True).stdout
    undefined = set(re.findall(r"\bU (cuembed_\w+)", syms))
    assert {"cuembed_forward", "cuembed_transpose", "cuembed_backward",
            "cuembed_compressed_grad_indices"} <= undefined, undefined
    assert "EmbeddingLookUpKernel" not in syms
    assert "EmbeddingBackwardKernel" not in syms
    assert "DeviceRadixSort::SortPairs" not in syms


@pytest.mark.gpu
@pytest.mark.parametrize("name", SUITES)
def test_reference_gtest_suite_passes_against_the_dropin(name):
    path = _need(name)
    r = subprocess.run([path, "--gtest_brief=1"], capture_output=True, text=True,
                       timeout=1500, cwd=CONF)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    m = re.search(r"\[  PASSED  \] (\d+) test", r.stdout)
    assert m and int(m.group(1)) > 0, tail
    assert "FAILED" not in r.stdout, tail
