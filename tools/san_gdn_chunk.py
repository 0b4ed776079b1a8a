#!/usr/bin/env python
"""Smallest run of the chunkwise Gated-Delta-Net kernels, for compute-sanitizer: a tiny Qwen3.5 model (2 + 4 heads of 128), one
130-row prefill (two full chunks + a 2-row tail) after a 20-row call that leaves a non-zero state, then one decode step.
    compute-sanitizer --tool memcheck python tools/san_gdn_chunk.py"""
import os
import sys
import time

t0 = time.time()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import crane_b200  # noqa: E402
from crane_b200 import synth  # noqa: E402

cfg = synth.TINY_QWEN3_5
m = crane_b200.Qwen3_5Model(cfg, device=0, max_seq_len=256, gdn="chunked")
m.load_checkpoint(synth.synth_checkpoint(cfg))
ids = synth.synth_token_ids(150, cfg["vocab_size"], "san")
m.forward_step(ids[:20], 0)
lg = m.forward_step(ids[20:], 20)
lg2 = m.forward_step([int(np.argmax(lg))], 150)
print(f"san_gdn_chunk: ok, logits finite {bool(np.isfinite(lg).all() and np.isfinite(lg2).all())}, {time.time() - t0:.1f} s", flush=True)
m.close()
