#!/bin/bash
# A/B of backward launch bounds / ownership knob on the GPU box.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log
show() { python -c "
import json;d=json.load(open('$1'));print('$2', {k:v['ms'] for k,v in d['stages'].items() if 'ms' in v}, d['ms_per_step'])"; }
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_own1.json 2>/dev/null; show gpurun_out/bench_own1.json own1
CUEMBED_BWD_OWN=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_own0.json 2>/dev/null; show gpurun_out/bench_own0.json own0
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Bwd' -c 6 --csv --log-file gpurun_out/launches_quick.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_l.log 2>&1
CUEMBED_NVCC_EXTRA="-DBWD_MINB=5" python -m cuembed_b200.build > /dev/null 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_minb5.json 2>/dev/null; show gpurun_out/bench_minb5.json minb5
