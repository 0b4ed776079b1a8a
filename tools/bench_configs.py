#!/usr/bin/env python
"""Device timings for the other BASELINE.json configs (the headline config 1 is bench.py):
  c0  Qwen3-0.6B bf16          128-token prefill + 128 greedy decode steps
  c2  Qwen3.5-0.8B hybrid      4096-token prefill + 512 decode steps
  c3  Qwen3-8B Q4_K_M-like     4 sequences x (128-token prompt, 256 batched decode steps)  [the per-GPU share of batch 32 on 8 GPUs]
  c4  Qwen3-TTS-12Hz-0.6B      135-position prefill + 125 greedy frames (1 talker + 16 code-predictor passes each)
Weights are cheap tiled random blocks (values irrelevant for timing); parity for every path is in tests/test_gpu_parity.py.
Prints one JSON line per config: achieved tok/s (or frames/s), HBM GB/s against the algorithmic bytes, prefill TFLOP/s."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crane_b200  # noqa: E402
from crane_b200 import synth  # noqa: E402

rng = np.random.default_rng(0)
BLOCK = synth.f32_to_bf16_bits(rng.standard_normal(1 << 22, dtype=np.float32) * 0.02)


def cheap(shape, kind):
    n = int(np.prod(shape))
    if kind in ("norm",):
        return np.ones(shape, np.float32)
    if kind in ("norm0", "bias", "dt_bias"):
        return np.zeros(shape, np.float32)
    if kind == "a_log":
        return np.full(shape, -2.0, np.float32)
    reps = (n + BLOCK.size - 1) // BLOCK.size
    return np.tile(BLOCK, reps)[:n].reshape(shape)


GGML_TYPE_ID = {"Q8_0": 8, "Q4_K": 12, "Q6_K": 14}


def fake_blocks(rows, K, qt):
    """Syntactically valid ggml blocks with random codes and small f16 scales (timing only: the values are irrelevant, and this tool
    must not touch oracle/).  Layouts: Q8_0 [f16 d | 32 i8]; Q4_K [f16 d | f16 dmin | 12 B scales | 128 B nibbles]; Q6_K [128 ql | 64 qh | 16 i8 | f16 d]."""
    small = np.frombuffer(np.float16(0.01).tobytes(), np.uint8)
    if qt == "Q8_0":
        b = rng.integers(0, 256, (rows, K // 32, 34), dtype=np.uint8)
        b[..., 0:2] = small
    elif qt == "Q4_K":
        b = rng.integers(0, 256, (rows, K // 256, 144), dtype=np.uint8)
        b[..., 0:2] = small
        b[..., 2:4] = small
    else:
        b = rng.integers(0, 256, (rows, K // 256, 210), dtype=np.uint8)
        b[..., 192:208] = rng.integers(0, 16, (rows, K // 256, 16), dtype=np.uint8)
        b[..., 208:210] = small
    return b.reshape(rows, -1)


def load_cheap(m, cfg, quant=None):
    qcache = {}
    for name, shape, kind in synth.tensor_specs(cfg):
        qt = None
        if quant and len(shape) == 2:
            qt = next((t for suf, t in quant.items() if name.endswith(suf)), None)
        if qt:
            key = (qt, shape[1])
            if key not in qcache:
                qcache[key] = fake_blocks(64, shape[1], qt)
            raw = np.tile(qcache[key], ((shape[0] + 63) // 64, 1))[:shape[0]]
            m.load_tensor_ggml(name, GGML_TYPE_ID[qt], shape, raw)
        else:
            m.load_tensor(name, cheap(shape, kind))
    m.finalize()


def dense_bytes_per_token(tc, ctx, wbytes=2.0, head_bytes=2.0, full_layers=None):
    H, I, L, V = tc["hidden_size"], tc["intermediate_size"], tc["num_hidden_layers"], tc["vocab_size"]
    nh, nkv, d = tc["num_attention_heads"], tc["num_key_value_heads"], tc["head_dim"]
    per_layer = ((nh + 2 * nkv) * d * H + H * nh * d + 2 * I * H + H * I) * wbytes
    return per_layer * L + V * H * head_bytes + ctx * 2 * nkv * d * 2 * (full_layers or L)


def c0():
    cfg = synth.QWEN3_0_6B
    m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=512)
    load_cheap(m, cfg)
    ids = synth.synth_token_ids(128, cfg["vocab_size"], "c0")
    best = None
    for _ in range(3):
        m.clear_kv_cache()
        tok = m.forward_step_argmax(ids, 0)
        m.decode_greedy(tok, 128, 128)
        t = m.last_timing()
        best = t if best is None or t["decode_ms"] < best["decode_ms"] else best
    us = best["decode_ms"] / 128 * 1e3
    by = dense_bytes_per_token(cfg, 192)
    m.close()
    return {"config": "c0 Qwen3-0.6B bf16, 128-tok prompt, 128 decode", "decode_tok_s": 1e6 / us, "decode_us_per_step": us,
            "decode_hbm_gbs": by / us / 1e3, "prefill_ms": best["prefill_ms"]}


def c2():
    cfg = synth.QWEN3_5_0_8B
    m = crane_b200.Qwen3_5Model(cfg, device=0, max_seq_len=4736)
    load_cheap(m, cfg)
    ids = synth.synth_token_ids(4096, cfg["vocab_size"], "c2")
    best = None
    for _ in range(2):
        m.clear_kv_cache()
        tok = m.forward_step_argmax(ids, 0)
        m.decode_greedy(tok, 4096, 512)
        t = m.last_timing()
        best = t if best is None or t["decode_ms"] < best["decode_ms"] else best
    us = best["decode_ms"] / 512 * 1e3
    H, I, L, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"], cfg["vocab_size"]
    full, gdn = 6, 18
    w_full = (8 * 256 * 2 + 2 * 2 * 256) * H + H * 8 * 256
    w_gdn = (6144 + 2048 + 32) * H + H * 2048
    by = ((w_full * full + w_gdn * gdn + 3 * I * H * L) + V * H) * 2 + gdn * 2 * 16 * 128 * 128 * 4 + 4352 * 2 * 2 * 256 * 2 * full
    fl = 2 * ((w_full * full + w_gdn * gdn + 3 * I * H * L)) * 4096 + 4 * 4096 * 2048 * 8 * 256 * full + 2 * V * H
    m.close()
    return {"config": "c2 Qwen3.5-0.8B hybrid bf16, 4096-tok prefill, 512 decode", "decode_tok_s": 1e6 / us, "decode_us_per_step": us,
            "decode_hbm_gbs": by / us / 1e3, "prefill_ms": best["prefill_ms"], "prefill_tflops": fl / best["prefill_ms"] / 1e9,
            "prefill_tok_s": 4096 / best["prefill_ms"] * 1e3}


def c3():
    cfg = synth.QWEN3_8B
    quant = {"q_proj.weight": "Q4_K", "k_proj.weight": "Q4_K", "v_proj.weight": "Q6_K", "o_proj.weight": "Q4_K", "gate_proj.weight": "Q4_K",
             "up_proj.weight": "Q4_K", "down_proj.weight": "Q6_K", "lm_head.weight": "Q6_K", "embed_tokens.weight": "Q4_K"}
    m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=512, max_batch=4)
    t0 = time.time()
    load_cheap(m, cfg, quant)
    print(f"c3 load {time.time() - t0:.1f}s", file=sys.stderr)
    prompts = [synth.synth_token_ids(128, cfg["vocab_size"], f"c3-{i}") for i in range(4)]
    seqs, first = [0], []
    m.seq_select(0)
    first.append(m.forward_step_argmax(prompts[0], 0))
    pre_ms = m.last_timing()["prefill_ms"]
    for p in prompts[1:]:
        s = m.seq_create()
        m.seq_select(s)
        first.append(m.forward_step_argmax(p, 0))
        seqs.append(s)
    steps = 256
    m.decode_batch(seqs, first, n_steps=8)            # warm-up
    t = time.perf_counter()
    m.decode_batch(seqs, [1, 2, 3, 4], n_steps=steps)
    wall = time.perf_counter() - t
    dm = m.last_timing()["decode_ms"]
    # bytes per batched step: Q4_K 0.5625 B/weight for q,k,o,gate,up; Q6_K 0.875 for v, down, head
    H, I, L, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"], cfg["vocab_size"]
    q4 = (32 * 128 * H + 8 * 128 * H + H * 32 * 128 + 2 * I * H) * 0.5625
    q6 = (8 * 128 * H + H * I) * 0.875
    by = (q4 + q6) * L + V * H * 0.875 + 4 * 264 * 2 * 8 * 128 * 2 * L
    us = dm / steps * 1e3
    # single-sequence decode for comparison
    m.seq_select(0)
    m.decode_greedy(5, m.kv_len(), 8)                 # warm-up (graph capture)
    m.decode_greedy(5, m.kv_len(), 64)
    us1 = m.last_timing()["decode_ms"] / 64 * 1e3
    m.close()
    return {"config": "c3 Qwen3-8B Q4_K_M-like, 4 seq/GPU x 128-tok prompt, 256 batched decode steps", "decode_tok_s": 4e6 / us,
            "decode_us_per_step": us, "decode_hbm_gbs": by / us / 1e3, "wall_tok_s": 4 * steps / wall, "prefill_ms_per_seq": pre_ms,
            "single_seq_decode_tok_s": 1e6 / us1, "single_seq_hbm_gbs": (by - 3 * 264 * 2 * 8 * 128 * 2 * L) / us1 / 1e3}


def c4():
    cfg = synth.QWEN3_TTS_0_6B
    m = crane_b200.Qwen3TTSModel(cfg, device=0, max_seq_len=1024)
    load_cheap(m, cfg)
    ids = synth.synth_token_ids(126, 150000, "c4")     # 9 + 126 prefill positions (10 s voice-clone-sized prefix)
    frames = 125
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        out = m.generate_codes(ids, frames, repetition_penalty=1.05)
        wall = time.perf_counter() - t0
        t = m.last_timing()
        t["wall"] = wall
        t["n"] = len(out)
        best = t if best is None or t["decode_ms"] < best["decode_ms"] else best
    n = max(best["decode_steps"], 1)
    us = best["decode_ms"] / n * 1e3
    tk, cp = cfg["talker_config"], cfg["talker_config"]["code_predictor_config"]
    by = dense_bytes_per_token(tk, 200) + 16 * (dense_bytes_per_token(dict(cp, vocab_size=2048), 8))
    m.close()
    return {"config": "c4 Qwen3-TTS-12Hz-0.6B, 135-position prefill + 125 greedy frames", "frames_per_s": 1e6 / us, "us_per_frame": us,
            "realtime_factor": (1e6 / us) / 12.5, "hbm_gbs": by / us / 1e3, "frames_launched": n, "frames_kept": best["n"],
            "prefill_ms": best["prefill_ms"]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c0", "c2", "c3", "c4"])
    args = ap.parse_args()
    for c in args.configs:
        try:
            print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in globals()[c]().items()}), flush=True)
        except Exception as e:      # keep going: one config must not hide the others
            print(json.dumps({"config": c, "error": repr(e)[:300]}), flush=True)
