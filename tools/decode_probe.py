#!/usr/bin/env python
"""Decode-path probe: text decoder of Qwen3-VL-2B geometry with cheap tiled random weights (values irrelevant for timing),
prefill of --ctx tokens then --steps decode steps; prints device time per step.  Used under ncu for the launch list."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crane_b200  # noqa: E402
from crane_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ctx", type=int, default=454)
ap.add_argument("--steps", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--model", default="vl2b", choices=["vl2b", "0.6b", "8b"])
args = ap.parse_args()

cfg = dict({"vl2b": synth.QWEN3_VL_2B["text_config"], "0.6b": synth.QWEN3_0_6B, "8b": synth.QWEN3_8B}[args.model], model_type="qwen3")
rng = np.random.default_rng(0)
block = synth.f32_to_bf16_bits(rng.standard_normal(1 << 22, dtype=np.float32) * 0.02)


def cheap(shape, kind):
    n = int(np.prod(shape))
    if kind == "norm":
        return np.ones(shape, np.float32)
    reps = (n + block.size - 1) // block.size
    return np.tile(block, reps)[:n].reshape(shape)


m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=2048)
t0 = time.time()
for name, shape, kind in synth.tensor_specs(cfg):
    m.load_tensor(name, cheap(shape, kind))
m.finalize()
print(f"load {time.time() - t0:.1f}s", file=sys.stderr)
ids = synth.synth_token_ids(args.ctx, cfg["vocab_size"], "probe")
for rep in range(args.reps):
    m.clear_kv_cache()
    tok = m.forward_step_argmax(ids, 0)
    m.decode_greedy(tok, args.ctx, args.steps)
    t = m.last_timing()
    print(f"rep {rep}: prefill {t['prefill_ms']:.3f} ms, decode {t['decode_ms'] / t['decode_steps'] * 1e3:.1f} us/step "
          f"({t['decode_steps'] / t['decode_ms'] * 1e3:.1f} tok/s)")
