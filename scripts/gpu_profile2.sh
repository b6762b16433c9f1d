#!/bin/bash
mkdir -p gpurun_out
for K in RadixPassKernel CompressScanKernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$K -s 0 -c 1 -f -o gpurun_out/prof0_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu0_$K.log 2>&1
  echo "full $K rc=$?"
done
