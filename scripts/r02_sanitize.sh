#!/bin/bash
# compute-sanitizer memcheck over the entry points added or rewritten in round 2
# (warp walker with one-warp CTAs, 12-key radix tiles, transpose_fixed, mapped
# forward, debug check, pool-and-push / select kernels on virtual ranks).
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_parity.py tests/test_sharded_p2p.py -m gpu -x -q \
  -k "kat or mapped or debug_check or transpose_fixed or hot_rows or long_runs or general_keys or virtual or pool or concat" \
  > gpurun_out/pytest_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/pytest_memcheck.log; grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -4 gpurun_out/memcheck.log
