#!/bin/bash
# ncu full capture of the backward walker for the default lib and a variant lib ($1)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:BwdWarpKernel -s 1 -c 1 -f -o gpurun_out/prof_BwdWarpKernel $B > gpurun_out/ncu_a.log 2>&1; echo "ncu default rc=$?"
if [ -n "$1" ]; then
CUEMBED_B200_LIB=$PWD/cuembed_b200/lib/$1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:BwdWarpKernel -s 1 -c 1 -f -o gpurun_out/prof_var_BwdWarpKernel $B > gpurun_out/ncu_b.log 2>&1; echo "ncu variant rc=$?"
fi
