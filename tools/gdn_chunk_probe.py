#!/usr/bin/env python
"""Prefill of a 4-layer Qwen3.5-0.8B-width model (3 Gated-Delta-Net layers + 1 full-attention layer, 16 + 16 heads of 128) with the
chunkwise or the token-by-token recurrence: CUDA-event time of the pass and, with --spans, the stage spans (crane_b200_prof_report).
    python tools/gdn_chunk_probe.py chunked|sequential [S=4096] [--spans]
Under `ncu --metrics gpu__time_duration.sum -k regex:gdn_` the same command lists the recurrence kernels one by one."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_configs  # noqa: E402
import crane_b200  # noqa: E402
from crane_b200 import synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "chunked"
S = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 4096
cfg = dict(synth.QWEN3_5_0_8B, num_hidden_layers=4, vocab_size=4096)
m = crane_b200.Qwen3_5Model(cfg, device=0, max_seq_len=S + 64, gdn=mode)
bench_configs.load_cheap(m, cfg)
ids = synth.synth_token_ids(S, cfg["vocab_size"], "probe")
times = []
for _ in range(4):
    m.clear_kv_cache()
    m.forward_step_argmax(ids, 0)
    times.append(m.last_timing()["prefill_ms"])
out = {"mode": mode, "S": S, "layers": 4, "gdn_layers": 3, "prefill_ms": round(min(times[1:]), 4), "all_ms": [round(t, 4) for t in times]}
if "--spans" in sys.argv:
    m.prof_enable(True)
    for _ in range(3):
        m.clear_kv_cache()
        m.forward_step_argmax(ids, 0)
    rep = m.prof_report()
    out["spans_device_ms"] = {k: round(v["device_ms"], 4) for k, v in rep["prefill"]["spans"].items()}
print(json.dumps(out), flush=True)
m.close()
