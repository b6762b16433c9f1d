#!/bin/bash
# BASELINE.json configs C3 (CSR, weighted sum / mean, fp32 + bf16, int64) and C4 (backward sweep:
# full vs compressed gradients, uniform vs power-law, width 64..512) through the
# manual_benchmark-compatible CLI.  Output: gpurun_out/configs_c3_c4.csv
mkdir -p gpurun_out
OUT=gpurun_out/configs_c3_c4.csv
rm -f $OUT
MB="python benchmarks/manual_benchmark.py --iterations 20 --enable_csv --csv_file $OUT --noenable_stderr"
C3="--num_categories 10000000 --embed_width 128 --batch_size 131072 --hotness 64 --alpha=1.15 --csr_input --use_int64_indices"
$MB $C3 --weighted_sum                       # fp32 weighted sum
$MB $C3 --combine_mode mean                  # fp32 mean
$MB $C3 --weighted_sum --bf16                # bf16 weighted sum
$MB $C3 --combine_mode mean --bf16           # bf16 mean
$MB $C3 --weighted_sum --combine_mode mean --forward_only   # weighted mean (forward only: no backward formula)
for alpha in 0.0 1.15; do
  for width in 64 128 256 512; do
    for grad in --compressed_grad --nocompressed_grad; do
      $MB --num_categories 10000000 --embed_width $width --batch_size 65536 --hotness 64 \
          --alpha=$alpha --half_embedding_type $grad
    done
  done
done
cat $OUT
