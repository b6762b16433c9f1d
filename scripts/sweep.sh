#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/sweep_$name.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stages']; print('$name', 'fwd %.4f tr %.4f bwd %.4f total %.4f ms'%(s['forward']['ms'],s['transpose']['ms'],s['backward']['ms'],d['ms_per_step']))"
}
run base X=1
run sort_nolookback CUEMBED_SORT_DEBUG=1
run sort_backoff CUEMBED_SORT_DEBUG=2
run sort_i16 CUEMBED_SORT_ITEMS=16
run sort_i16_backoff CUEMBED_SORT_ITEMS=16 CUEMBED_SORT_DEBUG=2
run sort_i16_nolb CUEMBED_SORT_ITEMS=16 CUEMBED_SORT_DEBUG=1
