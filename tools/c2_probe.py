#!/usr/bin/env python
"""Qwen3.5-0.8B with a reduced layer count, one 4096-token prefill: the workload for an ncu launch list of the hybrid prefill."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crane_b200
from crane_b200 import synth
from tools.bench_configs import load_cheap
L = int(os.environ.get("LAYERS", "8"))
cfg = dict(synth.QWEN3_5_0_8B, num_hidden_layers=L)
if "layer_types" in cfg:
    cfg["layer_types"] = cfg["layer_types"][:L]
m = crane_b200.Qwen3_5Model(cfg, device=0, max_seq_len=4352)
load_cheap(m, cfg)
ids = synth.synth_token_ids(4096, cfg["vocab_size"], "c2")
tok = m.forward_step_argmax(ids, 0)
print(m.last_timing())
