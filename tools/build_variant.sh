#!/bin/bash
# tools/build_variant.sh <name> "<extra nvcc flags>": builds build/variants/libubu_sw_<name>.so (select it with UBU_SW_LIB=...) for
# A/B runs inside ONE gpurun call (run-to-run and box-to-box noise is larger than most kernel deltas).
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/variants/$NAME; mkdir -p $OUT
for u in ubu_sw ubu_kmer ubu_pack; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $EXTRA -c -o $OUT/$u.o $ROOT/ubu_b200/csrc/$u.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/variants/libubu_sw_$NAME.so $OUT/ubu_sw.o $OUT/ubu_kmer.o $OUT/ubu_pack.o
rm -rf $OUT
echo $ROOT/variants/libubu_sw_$NAME.so
