#!/bin/bash
# A/B of an environment knob: bash scripts/gpu_ab.sh VAR val0 val1
mkdir -p gpurun_out
show() { python -c "
import json;d=json.load(open('$1'));print('$2', {k:v['ms'] for k,v in d['stages'].items() if 'ms' in v}, d['ms_per_step'])"; }
for V in $2 $3 $2 $3; do
  env $1=$V python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_ab_$V.json 2>/dev/null; show gpurun_out/bench_ab_$V.json "$1=$V"
done
