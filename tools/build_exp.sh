#!/bin/bash
# Tuning aid: builds roki-fd_b200/exp/librokifd_b200_<tag>.so = the product library with ONE kernel variant recompiled with
# extra flags (e.g. -DRKFD_SYNC_LEVEL=1).  Select it at run time with ROKIFD_B200_LIB=<path>.
#   tools/build_exp.sh <tag> "<extra nvcc flags>" [variant, default 128_0_0_5_4]
set -e
TAG=$1; EXTRA=$2; V=${3:-128_0_0_5_4}
cd "$(dirname "$0")/../roki-fd_b200/csrc"
mkdir -p build_exp ../exp
IFS=_ read B G R S M <<< "$V"
/usr/local/cuda/bin/nvcc $EXTRA -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v \
  -DRKFD_BLOCK=$B -DRKFD_GSCR=$G -DRKFD_RIGID=$R -DRKFD_SPEC=$S -DRKFD_MINB=$M -c rkfd_kernel_variant.cu -o build_exp/k_${TAG}.o 2> build_exp/ptxas_${TAG}.log
grep -E "Used [0-9]+ registers" build_exp/ptxas_${TAG}.log | head -1; grep -E "spill" build_exp/ptxas_${TAG}.log | head -1
OBJS=$(ls build/*.o | grep -v "rkfd_kernel_${V}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../exp/librokifd_b200_${TAG}.so $OBJS build_exp/k_${TAG}.o -cudart static
echo built ../exp/librokifd_b200_${TAG}.so
