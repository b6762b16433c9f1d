"""The drop-in C++ host templates (include/cuembed/include/*.cuh).

  * CPU: the reference's own caller translation units
    (utils/src/embedding_gpu_{forward,transpose,backward}.cu) compile UNCHANGED
    against the drop-in headers (needs /root/reference; skipped on the GPU box).
  * GPU: a native C++ program written against the reference API runs the
    reference's known-answer vectors through the templates.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "utils", "src")),
                    reason="reference tree not present")
@pytest.mark.parametrize("unit", ["forward", "transpose", "backward"])
def test_reference_callers_compile_unchanged(tmp_path, unit):
    src = os.path.join(REF, "utils", "src", f"embedding_gpu_{unit}.cu")
    obj = tmp_path / f"{unit}.o"
    cmd = ["nvcc", "-std=c++17", "-c", "-gencode", "arch=compute_100a,code=sm_100a",
           f"-I{ROOT}/include",            # cuembed/include/* -> the drop-in headers
           f"-I{REF}",                     # utils/include/*   -> the reference harness
           f"-I{REF}/third_party/abseil-cpp", src, "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    syms = subprocess.run(["nm", "-C", str(obj)], capture_output=True, text=True).stdout
    assert "U cuembed_" in syms  # the templates resolved to the C ABI


def test_dropin_program_builds():
    import __graft_entry__ as g
    exe = g.build_dropin_program()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_dropin_program_runs_kats():
    import __graft_entry__ as g
    exe = g.build_dropin_program()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failure(s)" in r.stdout
