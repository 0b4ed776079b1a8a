import os, sys
import numpy as np
os.environ["CRANE_B200_GEMM_PROF"] = "1"
sys.path.insert(0, "/root/repo")
import crane_b200
from crane_b200 import synth
rng = np.random.default_rng(0)
for name, M, N, K, mode in [("o", 454, 2048, 2048, crane_b200.EPI_RESID_F32), ("down", 454, 2048, 6144, crane_b200.EPI_RESID_F32), ("fc1", 784, 4096, 1024, crane_b200.EPI_GELU_ERF_BF16)]:
    a = rng.standard_normal((M, K), dtype=np.float32)
    hi = synth.bf16_round(a)
    w = synth.f32_to_bf16_bits(rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K))
    crane_b200.op_gemm(synth.f32_to_bf16_bits(a), w, mode, a_lo_bits=synth.f32_to_bf16_bits(a - hi))
