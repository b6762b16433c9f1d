#!/bin/bash
# A/B of backward build variants (scripts/build_variant.sh): tests with the w1 variant, then bench lines.
mkdir -p gpurun_out
L=cuembed_b200/lib
CUEMBED_B200_LIB=$PWD/$L/variant_w1.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -q -x \
   -k "backward or bwd or matrix or kat or optimizer or full_size" > gpurun_out/pytest_w1.log 2>&1; echo "pytest(w1) rc=$?"; tail -2 gpurun_out/pytest_w1.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
i=0
for v in "$@"; do
  env $v $B > gpurun_out/bench_w$i.json 2> gpurun_out/bench_w$i.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_w$i.json"))
    print("variant [$v]", d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("variant [$v] failed rc=$rc", e)
PY
  i=$((i+1))
done
