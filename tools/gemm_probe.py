#!/usr/bin/env python
"""tcgen05 GEMM timeline probe (run on the GPU box): for the prefill shapes of Qwen3-VL-2B, back-to-back event timing plus the
per-CTA timeline (setup / first operands / main loop / accumulator wait / epilogue) written by the kernel's globaltimer stamps.
  CRANE_B200_GEMM_CLUSTERS=0 python tools/gemm_probe.py      # without the multicast clusters"""
import os
import sys

import numpy as np

os.environ["CRANE_B200_GEMM_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crane_b200  # noqa: E402
from crane_b200 import synth  # noqa: E402

rng = np.random.default_rng(0)
SHAPES = [("qkv", 454, 4096, 2048, crane_b200.EPI_STORE_F32), ("o", 454, 2048, 2048, crane_b200.EPI_RESID_F32),
          ("gate_up", 454, 12288, 2048, crane_b200.EPI_SILU_MUL_BF16), ("down", 454, 2048, 6144, crane_b200.EPI_RESID_F32),
          ("vit_qkv", 784, 3072, 1024, crane_b200.EPI_STORE_F32), ("vit_fc1", 784, 4096, 1024, crane_b200.EPI_GELU_ERF_BF16),
          ("vit_fc2", 784, 1024, 4096, crane_b200.EPI_RESID_F32), ("big", 4096, 4096, 4096, crane_b200.EPI_STORE_F32)]
for name, M, N, K, mode in SHAPES:
    a = rng.standard_normal((M, K), dtype=np.float32)
    hi = synth.bf16_round(a)
    w = synth.f32_to_bf16_bits(rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K))
    for split in (False, True):
        print(f"--- {name} split={split}", file=sys.stderr, flush=True)
        crane_b200.op_gemm(synth.f32_to_bf16_bits(a), w, mode, a_lo_bits=synth.f32_to_bf16_bits(a - hi) if split else None)
