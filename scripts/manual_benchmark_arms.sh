#!/bin/bash
# Both arms of the same-box comparison from ONE unchanged benchmark source
# (/root/reference/benchmarks/manual_benchmark.cu, built by oracle/conformance.sh):
#   oracle/_ref/conformance/manual_benchmark      -> this library behind the drop-in headers
#   oracle/_ref/conformance/manual_benchmark_ref  -> the reference's own headers (sm_100 build)
# at the C2 shape (README.md:104 defaults of the reference sweep).  CSVs land in
# gpurun_out/manual_benchmark_{b200,ref}.csv.   usage: scripts/manual_benchmark_arms.sh [iterations]
IT=${1:-20}
CONF=oracle/_ref/conformance
mkdir -p gpurun_out
ARGS="--num_categories=10000000 --embed_width=256 --batch_size=65536 --hotness=64 --alpha=1.15 \
  --half_embedding_type=true --compressed_grad=true --skip_grad_init=true --iterations=$IT \
  --enable_csv=true --enable_stderr=false"
for arm in b200 ref; do
  exe=$CONF/manual_benchmark; [ $arm = ref ] && exe=$CONF/manual_benchmark_ref
  [ -x $exe ] || { echo "$exe missing"; continue; }
  d=$(mktemp -d); ( cd $d && timeout 900 $OLDPWD/$exe $ARGS > run.log 2>&1; echo "manual_benchmark[$arm] rc=$?" )
  cp $d/manual_benchmark_out.csv gpurun_out/manual_benchmark_$arm.csv 2>/dev/null
  tail -3 $d/run.log; cat gpurun_out/manual_benchmark_$arm.csv
done
