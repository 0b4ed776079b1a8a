#!/usr/bin/env python
"""A few decode steps of the bench model for an ncu launch list (CRANE_B200_GRAPHS=0): per-kernel durations of the decode chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crane_b200
from crane_b200 import synth
from tools.bench_configs import load_cheap
cfg = dict(synth.QWEN3_VL_2B["text_config"], num_hidden_layers=int(os.environ.get("LAYERS", "6")), model_type="qwen3")
m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=1024)
load_cheap(m, cfg)
ids = synth.synth_token_ids(454, cfg["vocab_size"], "p")
tok = m.forward_step_argmax(ids, 0)
m.decode_greedy(tok, 454, 8)
print(m.last_timing())
