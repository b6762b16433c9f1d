#!/bin/bash
# N = 2 run: two-GPU tests + sharded bench (weak-scaled C2) + C5.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_p2p.py tests/test_sharded.py -m gpu -q -x > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_n2.log
N=${1:-2}
run() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 10 --warmup 3 $2 > gpurun_out/bench_n${N}_$1.json 2> gpurun_out/bench_n${N}_$1.err
  echo "bench $1 rc=$?"; tail -c 2500 gpurun_out/bench_n${N}_$1.json; echo; grep -v "^W\|^\[W\|Warning\|warn" gpurun_out/bench_n${N}_$1.err | tail -5
}
run c2 "--no-e2e"
run c2_f32 "--no-e2e --partial-dtype f32"
run c5 "--no-e2e --workload C5"
