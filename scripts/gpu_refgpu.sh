#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_ref_gpu.py C2 20 > gpurun_out/ref_gpu_same_box.json 2> gpurun_out/ref_gpu_same_box.err; echo rc=$?; cat gpurun_out/ref_gpu_same_box.json; tail -3 gpurun_out/ref_gpu_same_box.err
