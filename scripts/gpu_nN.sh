#!/bin/bash
# N-GPU bench line: bash scripts/gpu_nN.sh N
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-300
