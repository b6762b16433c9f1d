#!/bin/bash
# A/B of the staged backward (CUEMBED_BWD_STAGE): backward tests + short bench per
# variant, then the gather-ceiling study with the cp.async variants.
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
i=0
for v in "$@"; do
  if [ -n "$v" ]; then
    env $v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_torch_ops.py -m gpu -q -x \
      -k "backward or bwd or matrix or kat or optimizer or full_size" > gpurun_out/pytest_s$i.log 2>&1; echo "[$v] pytest rc=$?"; tail -3 gpurun_out/pytest_s$i.log
  fi
  env $v $B > gpurun_out/bench_s$i.json 2> gpurun_out/bench_s$i.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_s$i.json"))
    print("variant [$v]", d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("variant [$v] failed rc=$rc", e)
PY
  i=$((i+1))
done
python scripts/bench_gather.py > gpurun_out/gather_study.json 2> gpurun_out/gather_study.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/gather_study.json"))
for s,res in d.items():
    print(s, {k:v.get("ms") for k,v in res.items()})
PY
