#!/bin/bash
# run ON the GPU box: sweep the quantised GEMV's pipeline parameters
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/qgemv_sweep.txt
: > $out
for v in "" "-DQG_R_DEF=4 -DQG_DEPTH_DEF=3 -DQG_W1_DEF=16 -DQG_W4_DEF=12" "-DQG_R_DEF=4 -DQG_DEPTH_DEF=2 -DQG_W1_DEF=20 -DQG_W4_DEF=14"; do
  echo "=== $v" >> $out
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 $v tools/qgemv_bench.cu -o /tmp/qb 2>> $out && timeout 120 /tmp/qb >> $out 2>&1
done
cat $out
