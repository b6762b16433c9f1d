#!/bin/bash
# The reference's sweep grid (benchmarks/sweep_parameters.sh:21-31 of NVIDIA/cuEmbed)
# on this library: 3 alphas x 2 table sizes x 2 widths x 3 batches x 3 hotness values,
# fp32 / int32, results appended to manual_benchmark_out.csv.
#   bash benchmarks/sweep_parameters.sh [iterations] [extra flags...]
cd "$(dirname "$0")/.."
iterations=${1:-1000}
shift
rm -f manual_benchmark_out.csv
for alpha in 0.0 1.05 1.15; do
  for num_categories in 1000000 10000000; do
    for embed_width in 32 128; do
      for batch in 1024 32768 131072; do
        for hotness in 1 16 64; do
          python benchmarks/manual_benchmark.py --num_categories "${num_categories}" \
            --embed_width "${embed_width}" --batch_size "${batch}" --alpha=${alpha} \
            --hotness="${hotness}" --iterations="${iterations}" --enable_csv --noenable_stderr "$@"
        done
      done
    done
  done
done
