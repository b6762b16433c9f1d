#!/bin/bash
# Full evidence run: GPU tests, bench line, reference arm, ncu launch list, ncu full captures.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/gpu_launches.sh
for K in FwdPoolKernel BwdSegReduceKernel RadixPassKernel; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_$K.log 2>&1
  echo "full $K rc=$?"
done
