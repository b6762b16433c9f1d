#!/bin/bash
# compute-sanitizer memcheck over a subset of the GPU parity tests.
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "kat or fused or multi or hot_rows or empty_and_ragged or wide_and_odd or long_runs" > gpurun_out/pytest_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/pytest_memcheck.log; grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log
