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
    """Every test of the suite runs.  The 1026-case against-CPU suite spends its
    time in the reference's single-threaded CPU implementation (6 minutes in one
    process), so it is split with gtest's own sharding (GTEST_TOTAL_SHARDS /
    GTEST_SHARD_INDEX: each test runs in exactly one shard) over the host cores."""
    path = _need(name)
    shards = min(8, os.cpu_count() or 1) if name == "test_embedding_against_cpu" else 1
    procs = []
    for i in range(shards):
        env = dict(os.environ, GTEST_TOTAL_SHARDS=str(shards), GTEST_SHARD_INDEX=str(i))
        procs.append(subprocess.Popen([path, "--gtest_brief=1"], stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True, cwd=CONF, env=env))
    passed = 0
    for i, pr in enumerate(procs):
        try:
            out, _ = pr.communicate(timeout=1500)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        tail = out[-3000:]
        assert pr.returncode == 0, f"shard {i}/{shards}: {tail}"
        assert "FAILED" not in out, tail
        m = re.search(r"\[  PASSED  \] (\d+) test", out)
        assert m, tail
        passed += int(m.group(1))
    assert passed > 0
    # none lost to the split: the shards together ran every test the binary lists
    listed = subprocess.run([path, "--gtest_list_tests"], capture_output=True, text=True,
                            cwd=CONF).stdout
    assert passed == sum(1 for ln in listed.splitlines() if ln.startswith("  ")), passed
