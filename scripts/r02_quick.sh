#!/bin/bash
# Quick A/B run: backward-related GPU tests + short bench lines for env variants.
# usage: scripts/r02_quick.sh "VAR=val ..." "VAR=val ..." (each arg = one variant; "" = default)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_torch_ops.py tests/test_manual_benchmark.py -m gpu -q -x \
   -k "backward or bwd or hot or matrix or kat or optimizer or full_size or pipeline" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_quick.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
i=0
for v in "$@"; do
  env $v $B > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q$i.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q$i.json"))
    print("variant [$v]", d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("variant [$v] failed rc=$rc", e)
PY
  i=$((i+1))
done
bash scripts/gpu_launches.sh
