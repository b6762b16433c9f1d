#!/bin/bash
# Two-GPU check: sharded tests + N=2 bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_p2p.py tests/test_sharded.py -m gpu -x -q > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-700
