#!/bin/bash
# First-contact script for the GPU box: every stage in its own process so one CUDA fault does not hide the rest.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/smi.txt
nproc | tee gpurun_out/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)" | tee -a gpurun_out/nproc.txt
for t in "test_gemm_store_f32" "test_gemm_epilogues" "test_tiny_qwen3_against_hf_fixture" "test_chunked" "test_on_device" "test_forward_embeds" "test_tiny_qwen3_vl" "test_vl_generate"; do
  echo "=== $t"
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "$t" 2>&1 | tail -40
done 2>&1 | tee gpurun_out/bringup.log
