#!/bin/bash
# ncu evidence for the C2 step: launch list (durations) + one full capture of
# each of the three main kernels.  Run under gpurun; outputs in gpurun_out/.
mkdir -p gpurun_out
KREGEX='regex:FwdPool|FwdConcat|Radix|Compress|RowIds|BwdSeg|BwdFix'
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k "$KREGEX" -s 30 -c 20 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; tail -25 gpurun_out/launches.csv | cut -c1-220
for K in FwdPoolKernel BwdSegReduceKernel RadixPassKernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$K.log 2>&1
  echo "full $K rc=$?"
done
ls -la gpurun_out/
