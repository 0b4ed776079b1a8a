#!/bin/bash
# ncu evidence for the bench command (one GPU).  Usage: tools/profile.sh <round-tag>
cd "$(dirname "$0")/.."
TAG=${1:-r01}
mkdir -p gpurun_out
export CRANE_B200_GRAPHS=0
# 1. every launch of one request with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
# 2. full-set capture of the dominant decode kernel (gemv) and of the tcgen05 GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 300 -c 6 \
    -o gpurun_out/prof_gemv_${TAG} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_gemv_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 230 -c 4 \
    -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_gemm_${TAG}.log 2>&1
ls -la gpurun_out
