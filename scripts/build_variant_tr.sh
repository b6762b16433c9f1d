#!/bin/bash
# A differently tuned build of transforms.cu linked with the other objects of the
# product library: scripts/build_variant_tr.sh NAME "-DFLAG=.."  -> cuembed_b200/lib/variant_NAME.so
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME=$1; FLAGS=$2
L=$ROOT/cuembed_b200/lib
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $FLAGS \
  -c $ROOT/cuembed_b200/csrc/transforms.cu -o /tmp/transforms_$NAME.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $L/variant_$NAME.so $(ls $L/*.o | grep -v transforms.o) /tmp/transforms_$NAME.o -ldl
echo built $L/variant_$NAME.so
