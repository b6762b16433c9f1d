#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json;d=json.load(open('$1'));print('$2', {k:v['ms'] for k,v in d['stages'].items() if 'ms' in v}, d['ms_per_step'], d['gpu_launches'], d['config']['launch'])"; }
for i in 1 2; do
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; show gpurun_out/bench_graph.json graphs; tail -2 gpurun_out/bench_graph.err
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-graphs > gpurun_out/bench_nograph.json 2>/dev/null; show gpurun_out/bench_nograph.json direct
done
