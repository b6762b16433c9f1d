#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_size.py -m gpu -x -q --durations=5 > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_full.log
