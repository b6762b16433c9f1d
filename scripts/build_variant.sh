#!/bin/bash
# A differently tuned build of backward.cu linked with the other objects of the
# product library: scripts/build_variant.sh NAME "-DFLAG=.. ..."
# -> cuembed_b200/lib/variant_NAME.so (select with CUEMBED_B200_LIB=<path>).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME=$1; FLAGS=$2
L=$ROOT/cuembed_b200/lib
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $FLAGS \
  -c $ROOT/cuembed_b200/csrc/backward.cu -o /tmp/backward_$NAME.o
OBJS=$(ls $L/*.o | grep -v backward.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $L/variant_$NAME.so $OBJS /tmp/backward_$NAME.o -ldl
echo built $L/variant_$NAME.so
