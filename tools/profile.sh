#!/bin/bash
# ncu evidence for the bench command (one GPU).  Usage: tools/profile.sh <round-tag>     (run ON the GPU box, ~6 GPU-minutes)
cd "$(dirname "$0")/.."
TAG=${1:-r01}
mkdir -p gpurun_out
export CRANE_B200_GRAPHS=0          # every kernel of a decode step is its own launch under the profiler
BENCH="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
# 1. every launch of one request with its device time (cold-cache, serialised: compare SHARES) -- prefill + the first decode steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
# 2. full-set captures: six consecutive GEMVs of a decode step, the lm_head GEMV, the decode attention, tcgen05 GEMMs, flash prefill
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 300 -c 6 \
    -o gpurun_out/prof_gemv_${TAG} -f $BENCH > gpurun_out/ncu_gemv_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemv_kernel<1, 3" -s 2 -c 1 \
    -o gpurun_out/prof_lmhead_${TAG} -f $BENCH > gpurun_out/ncu_lmhead_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 60 -c 2 \
    -o gpurun_out/prof_attn_${TAG} -f $BENCH > gpurun_out/ncu_attn_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 230 -c 5 \
    -o gpurun_out/prof_gemm_${TAG} -f $BENCH > gpurun_out/ncu_gemm_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_prefill_kernel -s 30 -c 2 \
    -o gpurun_out/prof_flash_${TAG} -f $BENCH > gpurun_out/ncu_flash_${TAG}.log 2>&1
ls -la gpurun_out | tail -20
