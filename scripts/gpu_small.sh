#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_torch_ops.py tests/test_dropin_cpp.py tests/test_manual_benchmark.py -m gpu -x -q > gpurun_out/pytest_small.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_small.log
