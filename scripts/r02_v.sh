#!/bin/bash
# usage: scripts/r02_v.sh "<pytest -k expr>" "ENV=.. --flags" ...   (each further arg = env assignments and bench flags)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_cabi.py tests/test_manual_benchmark.py -m gpu -q -x \
   -k "$1" > gpurun_out/pytest_v.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_v.log
shift
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
i=0
for v in "$@"; do
  envs=""; flags=""
  for tok in $v; do case "$tok" in *=*) envs="$envs $tok";; *) flags="$flags $tok";; esac; done
  env $envs $B $flags > gpurun_out/bench_v$i.json 2> gpurun_out/bench_v$i.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v$i.json"))
    print("variant [$v]", d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("variant [$v] failed rc=$rc", e)
PY
  i=$((i+1))
done
