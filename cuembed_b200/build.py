"""Builds cuembed_b200/lib/libcuembed_b200.so from csrc/*.cu with nvcc for
sm_100a only (cross-compiles without a GPU).  In-tree so the .so travels with
the repository snapshot; `python -m cuembed_b200.build` or
__graft_entry__.build() run it."""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcuembed_b200.so")
SOURCES = ["c_api.cu", "forward.cu", "transforms.cu", "backward.cu", "sharded.cu",
           "sharded_p2p.cu", "microbench.cu", "forward_hot.cu", "debug.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-Xcompiler", "-fPIC",
    # tuning experiments only, e.g. CUEMBED_NVCC_EXTRA="-DBWD_MINB=6"
    *os.environ.get("CUEMBED_NVCC_EXTRA", "").split(),
]


def _deps() -> list[str]:
    root_inc = os.path.join(os.path.dirname(_HERE), "include", "cuembed_b200.h")
    return [root_inc] + [
        os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
        if f.endswith((".cuh", ".h"))
    ]


def _digest(paths: list[str]) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile_one(src: str, force: bool) -> str:
    path = os.path.join(CSRC, src)
    obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
    stamp = obj + ".sha"
    dig = _digest([path] + _deps())
    if not force and os.path.exists(obj) and os.path.exists(stamp):
        if open(stamp).read() == dig:
            return obj
    cmd = ["nvcc", *NVCC_FLAGS, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(sources)) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, force), sources))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", LIB_PATH, *objs, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
