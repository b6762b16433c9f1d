#!/bin/bash
# Quick check: backward parity tests + bench stage times + launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${1:-backward or against_cpu}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_quick.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print({k:v['ms'] for k,v in d['stages'].items() if 'ms' in v}, d['ms_per_step'], d.get('extras'))"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Bwd|Fwd|Radix|Compress|RowIds' -c 26 --csv --log-file gpurun_out/launches_quick.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
