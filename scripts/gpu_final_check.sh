#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_chk.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_chk.json'));print({k:v['ms'] for k,v in d['stages'].items() if 'ms' in v}, d['ms_per_step'])"
