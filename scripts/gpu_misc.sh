#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_multi_table.py > gpurun_out/multi_table.txt 2>&1; cat gpurun_out/multi_table.txt | tail -6
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; python -c "
import json;d=json.load(open('gpurun_out/bench_extras.json'));print(d['e2e']);print(d['extras'])"
bash scripts/configs_c3_c4.sh > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log
