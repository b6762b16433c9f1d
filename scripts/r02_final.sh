#!/bin/bash
# Evidence run of round 2: bench line (all legs), reference arm, smoke, ncu launch
# list, one `ncu --set full` capture of each of the three main kernels.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/gpu_launches.sh
for K in FwdPoolKernel BwdWarpKernel RadixPassKernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_$K.log 2>&1
  echo "full $K rc=$?"
done
