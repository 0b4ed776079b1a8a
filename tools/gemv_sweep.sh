#!/bin/bash
# run ON the GPU box: sweep the bf16 GEMV's (warps, depth, CTAs/SM)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/gemv_sweep4.txt
: > $out
for cfg in "32 3 1" "28 3 1" "24 4 1" "32 2 1"; do
  set -- $cfg
  echo "=== warps=$1 depth=$2 minb=$3" >> $out
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGV_WARPS=$1 -DGV_DEPTH=$2 -DGV_MINB=$3 tools/gemv_bench.cu -o /tmp/gb 2>> $out && timeout 120 /tmp/gb >> $out 2>&1
done
cat $out
