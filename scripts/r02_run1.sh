#!/bin/bash
# Round-2 evidence run: GPU tests (incl. the reference's own gtest suites against the
# drop-in), bench line with the reference-GPU leg and measured gather ceilings,
# reference arm, both manual_benchmark arms, ncu launch list + full captures.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
bash scripts/manual_benchmark_arms.sh 20
bash scripts/gpu_launches.sh
for K in FwdPoolKernel BwdSegReduceKernel RadixPassKernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_$K.log 2>&1
  echo "full $K rc=$?"
done
