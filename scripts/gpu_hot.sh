#!/bin/bash
# Backward experiments: parity tests, A/B bench (hot-row path off / on), launch lists.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward or against_cpu" > gpurun_out/pytest_hot.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_hot.log
for H in 0 1; do
CUEMBED_BWD_HOT=$H python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_hot$H.json 2> gpurun_out/bench_hot$H.err; echo "bench hot=$H rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_hot$H.json'));print({k:v['ms'] for k,v in d['stages'].items() if 'ms' in v}, d['ms_per_step'])"
CUEMBED_BWD_HOT=$H ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Bwd|Fwd|Radix|Compress|RowIds' -c 40 --csv --log-file gpurun_out/launches_hot$H.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
done
