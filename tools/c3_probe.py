#!/usr/bin/env python
"""Few-layer Qwen3-8B Q4_K_M-like model, a handful of single and batched decode steps: the workload for an ncu launch list
of the quantised decode kernels (run with CRANE_B200_GRAPHS=0 so every kernel is a separate launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crane_b200  # noqa: E402
from crane_b200 import synth  # noqa: E402
from tools.bench_configs import load_cheap  # noqa: E402

cfg = dict(synth.QWEN3_8B, num_hidden_layers=int(os.environ.get("LAYERS", "4")))
quant = {"q_proj.weight": "Q4_K", "k_proj.weight": "Q4_K", "v_proj.weight": "Q6_K", "o_proj.weight": "Q4_K", "gate_proj.weight": "Q4_K",
         "up_proj.weight": "Q4_K", "down_proj.weight": "Q6_K", "lm_head.weight": "Q6_K", "embed_tokens.weight": "Q4_K"}
m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=512, max_batch=4)
load_cheap(m, cfg, quant)
seqs = [0]
m.seq_select(0)
m.forward_step_argmax(synth.synth_token_ids(128, cfg["vocab_size"], "p0"), 0)
for i in range(3):
    s = m.seq_create()
    m.seq_select(s)
    m.forward_step_argmax(synth.synth_token_ids(128, cfg["vocab_size"], f"p{i + 1}"), 0)
    seqs.append(s)
m.decode_batch(seqs, [1, 2, 3, 4], n_steps=3)
m.seq_select(0)
m.decode_greedy(5, m.kv_len(), 3)
print("ok", m.last_timing())
