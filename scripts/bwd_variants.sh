#!/bin/bash
# Backward tuning: rebuild with different register budgets / chunk lengths and time the stages.
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/var_$name.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stages']; print('$name', 'fwd %.4f tr %.4f bwd %.4f total %.4f ms'%(s['forward']['ms'],s['transpose']['ms'],s['backward']['ms'],d['ms_per_step']))"
}
for MINB in 8 7 6; do
  CUEMBED_NVCC_EXTRA="-DBWD_MINB=$MINB" python -m cuembed_b200.build > /dev/null 2>&1
  export CUEMBED_NVCC_EXTRA="-DBWD_MINB=$MINB"
  run minb${MINB} X=1
  run minb${MINB}_r4 CUEMBED_BWD_ROUNDS=4
  run minb${MINB}_r16 CUEMBED_BWD_ROUNDS=16
  # (the unroll-4 variant of round 1 is gone: the walker is instantiated with 8 row loads in flight)
done
