#!/bin/bash
# Transpose-focused run: transform tests + short bench lines for flag variants.
# usage: scripts/r02_t.sh "flags..." "flags..."  (each arg = extra bench flags; "" = default)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_cabi.py tests/test_manual_benchmark.py -m gpu -q -x \
   -k "${TEST_K:-transpose or compress or matrix or kat or full_size or pipeline or cabi}" > gpurun_out/pytest_t.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_t.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
i=0
for v in "$@"; do
  $B $v > gpurun_out/bench_t$i.json 2> gpurun_out/bench_t$i.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_t$i.json"))
    print("variant [$v]", d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("variant [$v] failed rc=$rc", e)
PY
  i=$((i+1))
done
bash scripts/gpu_launches.sh
