"""CPU oracle for the dense Qwen3 decoder forward pass -- TEST INFRASTRUCTURE ONLY.

A restatement, in torch-CPU float32, of the arithmetic the reference's Candle CPU path
runs for `Qwen3Model::forward` / `forward_embeds`.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the product
(`crane_b200/`) never does.

The reference (Rust + Candle 0.11) cannot be built here (no cargo; candle is an un-vendored
crates.io dependency, Cargo.lock git-ignored), so its primitives (rms_norm, rope_thd,
softmax, flash_attn, matmul) are restated from their published definitions and pinned by
  * the reference's own literal known answers (tests/test_oracle_golden.py), and
  * HF transformers 5.5.0 `Qwen3ForCausalLM`, which the reference names as its ground truth
    (README.md:401-404), on random-init tiny configs (oracle/make_golden.py).

Candle's CPU backend has no bf16 matmul: the reference's CPU path computes in f32 on
weights up-cast from the bf16 safetensors (crane-core/src/models/qwen3/modeling.rs:1629-1632,
crane-serve/src/lib.rs:432-457).  So: weights = bf16-representable f32, arithmetic = f32.

Reference call path restated here (all crane-core/src/models/):
  Qwen3Model::forward / forward_embeds / decode ........ qwen3/modeling.rs:942-1036
  DecoderLayer::forward ................................. qwen3/modeling.rs:698-716
  Attention::forward (CPU flash-attn branches) .......... qwen3/modeling.rs:307-456
  Mlp::forward (merged gate_up, CPU branch) ............. qwen3/modeling.rs:608-642
  RotaryEmbedding::new / forward ........................ modules/rotary.rs:29-46,86-90
  update_kv_cache ....................................... modules/kv_cache.rs:38-101
"""
from __future__ import annotations

import math

import numpy as np
import torch

from crane_b200.synth import head_dim as _head_dim, text_config


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """candle_nn::ops::rms_norm: x * rsqrt(mean(x^2) + eps) * w, f32 accumulation
    (call sites qwen3/modeling.rs:208-217,660-669,784)."""
    var = x.pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(var + eps) * w


def rope_tables(dim: int, max_pos: int, theta: float):
    """RotaryEmbedding::new (modules/rotary.rs:29-46): inv_freq in f64 -> f32,
    freqs = pos_f32 (x) inv_freq as an f32 product, cos/sin f32 tables [max_pos, dim/2]."""
    inv = np.array([1.0 / (theta ** (i / dim)) for i in range(0, dim, 2)], dtype=np.float64).astype(np.float32)
    pos = np.arange(max_pos, dtype=np.float32)
    freqs = (pos[:, None] * inv[None, :]).astype(np.float32)
    return torch.from_numpy(np.cos(freqs)), torch.from_numpy(np.sin(freqs))


def rope_half(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """candle rope_thd / rope (non-interleaved, "NeoX" half-split): x [S, H, D],
    cos/sin [S, D/2]; (x1, x2) -> (x1 c - x2 s, x1 s + x2 c)  (qwen3/modeling.rs:358-359)."""
    d2 = x.shape[-1] // 2
    x1, x2 = x[..., :d2], x[..., d2:]
    c, s = cos[:, None, :], sin[:, None, :]
    return torch.cat([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1)


def causal_attention(q, k, v, kv_offset: int, scale: float):
    """candle_nn cpu flash_attn with AttnMask::Causal{kv_offset} (qwen3/modeling.rs:422-456)
    / AttnMask::None for decode (:380-420): softmax(q k^T * scale) v, f32 accumulation, GQA by
    integer division of the head index.  q [S, nh, d]; k, v [T, nkv, d] -> [S, nh*d]."""
    S, nh, d = q.shape
    T, nkv, _ = k.shape
    n_rep = nh // nkv
    qh = q.permute(1, 0, 2)                               # [nh, S, d]
    kh = k.permute(1, 0, 2).repeat_interleave(n_rep, 0)   # [nh, T, d]
    vh = v.permute(1, 0, 2).repeat_interleave(n_rep, 0)
    scores = torch.matmul(qh, kh.transpose(1, 2)) * scale  # [nh, S, T]
    i = torch.arange(S)[:, None]
    j = torch.arange(T)[None, :]
    scores = scores.masked_fill(j > (i + kv_offset), float("-inf"))
    p = torch.softmax(scores, dim=-1)
    o = torch.matmul(p, vh)                               # [nh, S, d]
    return o.permute(1, 0, 2).reshape(S, nh * d)


def silu(x):
    """candle Activation::Silu: x / (1 + exp(-x))."""
    return x * torch.sigmoid(x)


def quant_kv_per_token(x: torch.Tensor, bits: int) -> torch.Tensor:
    """`quantize_per_token` then `dequantize_per_token` (qwen3_5/kv_cache.rs:238-272): per (position, head) symmetric codes,
    scale = amax / qmax + 1e-8 (qmax 127 or 7), code = round(x / scale) (f32 `round`: half away from zero), stored with an offset
    of 128 / 8 (and two nibbles per byte for 4 bits -- storage only); what attention sees is code * scale.  x [..., D] f32."""
    qmax = float((1 << (bits - 1)) - 1)
    x = x.float()
    scale = x.abs().amax(dim=-1, keepdim=True) * (1.0 / qmax) + 1e-8
    q = x / scale
    q = torch.sign(q) * torch.floor(q.abs() + 0.5)
    return q * scale


class Qwen3Oracle:
    """Stateful (KV-cached) single-sequence forward, mirroring `Qwen3Model`."""

    def __init__(self, cfg: dict, weights: dict, prefix: str = "model.", max_pos: int | None = None, quantised: dict | None = None,
                 kv_bits: int = 0):
        """`quantised`: full tensor name -> (raw ggml blocks [rows, row_bytes] uint8, "Q4_K" | "Q6_K" | "Q8_0").  Those linears run
        through candle's CPU `QMatMul` semantics (ops/linear.rs:23-48): activations quantised to Q8_K / Q8_0 blocks, ggml integer
        dots (oracle/ggml_quant.py `qmatmul`); their entry in `weights` is not used for the product."""
        tc = text_config(cfg)
        self.tc = tc
        self.H = tc["hidden_size"]
        self.nh = tc["num_attention_heads"]
        self.nkv = tc["num_key_value_heads"]
        self.d = _head_dim(tc)
        self.L = tc["num_hidden_layers"]
        self.V = tc["vocab_size"]
        self.eps = float(tc.get("rms_norm_eps", 1e-6))
        self.theta = float(tc.get("rope_theta", 1_000_000.0))
        self.prefix = prefix
        self.q = dict(quantised or {})
        self.kv_bits = kv_bits         # 0: lossless cache; 8 / 4: QuantKvCache (qwen3_5/kv_cache.rs:209-342), see quant_kv_per_token
        self.w = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).float() for k, v in weights.items()}
        tied = cfg.get("tie_word_embeddings", tc.get("tie_word_embeddings", True))
        self.lm_head = self.w[prefix + "embed_tokens.weight"] if tied or "lm_head.weight" not in self.w \
            else self.w["lm_head.weight"]
        self.max_pos = max_pos or min(tc.get("max_position_embeddings", 4096), 8192)
        self.cos, self.sin = rope_tables(self.d, self.max_pos, self.theta)
        self.clear_kv_cache()
        self.last_hidden_states = None

    # -- KV cache (modules/kv_cache.rs:38-101: append in place, views over [0, T)) --------
    def clear_kv_cache(self):
        self.k_cache = [None] * self.L
        self.v_cache = [None] * self.L

    def kv_len(self) -> int:
        return 0 if self.k_cache[0] is None else self.k_cache[0].shape[0]

    def _p(self, i, name):
        return self.w[f"{self.prefix}layers.{i}.{name}"]

    def _linear(self, full_name: str, x: torch.Tensor, w: torch.Tensor | None = None) -> torch.Tensor:
        """x @ W^T, or LinearLayer::Quantized (ops/linear.rs:23-48: x -> f32, QMatMul::forward, cast back) when `full_name` is quantised."""
        if full_name in self.q:
            from . import ggml_quant as gq
            raw, qt = self.q[full_name]
            x2 = x.reshape(-1, x.shape[-1]).numpy()
            return torch.from_numpy(gq.qmatmul(x2, raw, qt)).reshape(*x.shape[:-1], -1)
        return x @ (self.w[full_name] if w is None else w).T

    def _lin(self, i, name, x):
        return self._linear(f"{self.prefix}layers.{i}.{name}", x)

    def embed(self, ids) -> torch.Tensor:
        ids = torch.as_tensor(np.asarray(ids, dtype=np.int64))
        return self.w[self.prefix + "embed_tokens.weight"][ids]

    # -- one decoder layer (qwen3/modeling.rs:698-716) ---------------------------------------
    def _layer(self, i, x, cos, sin, kv_offset):
        S = x.shape[0]
        h = rms_norm(x, self._p(i, "input_layernorm.weight"), self.eps)
        q = self._lin(i, "self_attn.q_proj.weight", h).view(S, self.nh, self.d)
        k = self._lin(i, "self_attn.k_proj.weight", h).view(S, self.nkv, self.d)
        v = self._lin(i, "self_attn.v_proj.weight", h).view(S, self.nkv, self.d)
        q = rms_norm(q, self._p(i, "self_attn.q_norm.weight"), self.eps)   # QK-norm BEFORE RoPE
        k = rms_norm(k, self._p(i, "self_attn.k_norm.weight"), self.eps)
        q = rope_half(q, cos, sin)
        k = rope_half(k, cos, sin)
        if self.kv_bits:
            k, v = quant_kv_per_token(k, self.kv_bits), quant_kv_per_token(v, self.kv_bits)
        if self.k_cache[i] is None:
            self.k_cache[i], self.v_cache[i] = k, v
        else:
            self.k_cache[i] = torch.cat([self.k_cache[i], k], 0)
            self.v_cache[i] = torch.cat([self.v_cache[i], v], 0)
        a = causal_attention(q, self.k_cache[i], self.v_cache[i], kv_offset, 1.0 / math.sqrt(self.d))
        x = x + self._lin(i, "self_attn.o_proj.weight", a)
        h = rms_norm(x, self._p(i, "post_attention_layernorm.weight"), self.eps)
        g = self._lin(i, "mlp.gate_proj.weight", h)
        u = self._lin(i, "mlp.up_proj.weight", h)
        x = x + self._lin(i, "mlp.down_proj.weight", silu(g) * u)
        return x

    def _cos_sin(self, start_pos, S):
        return self.cos[start_pos:start_pos + S], self.sin[start_pos:start_pos + S]

    def forward_embeds(self, x: torch.Tensor, start_pos: int, cos_sin=None, after_layer=None) -> torch.Tensor:
        """Qwen3Model::forward_embeds / decode (qwen3/modeling.rs:964-1036).  x [S, H] f32.
        Returns logits of the LAST position, [V]."""
        S = x.shape[0]
        assert start_pos == self.kv_len(), "start_pos must equal the cached length"
        cos, sin = cos_sin if cos_sin is not None else self._cos_sin(start_pos, S)
        for i in range(self.L):
            x = self._layer(i, x, cos, sin, start_pos)
            if after_layer is not None:
                x = after_layer(i, x)
        x = rms_norm(x, self.w[self.prefix + "norm.weight"], self.eps)
        self.last_hidden_states = x
        head = "lm_head.weight" if self.lm_head is self.w.get("lm_head.weight") else self.prefix + "embed_tokens.weight"
        return self._linear(head, x[-1:], self.lm_head)[0]

    def forward(self, ids, start_pos: int) -> torch.Tensor:
        """Qwen3Model::forward (qwen3/modeling.rs:942-953)."""
        return self.forward_embeds(self.embed(ids), start_pos)

    # -- Model::generate greedy branch (qwen3/model.rs:275-349): temp=None => ArgMax ---------
    def generate_greedy(self, prompt_ids, max_new_tokens: int, eos=()):
        self.clear_kv_cache()
        toks = list(int(t) for t in prompt_ids)
        out, margins = [], []
        for index in range(max_new_tokens):
            ctx = toks if index == 0 else toks[-1:]
            start = len(toks) - len(ctx)
            logits = self.forward(ctx, start)
            top2 = torch.topk(logits, 2)
            nxt = int(argmax_first(logits))
            margins.append(float(top2.values[0] - top2.values[1]))
            toks.append(nxt)
            out.append(nxt)
            if nxt in eos:
                break
        return out, margins


def argmax_first(logits: torch.Tensor) -> int:
    """Greedy rule of the oracle: lowest index among maxima."""
    m = logits.max()
    return int(torch.nonzero(logits == m)[0, 0])


def apply_repeat_penalty(logits: np.ndarray, penalty: float, context) -> np.ndarray:
    """models/utils.rs:25-44: once per distinct token, logit>=0 ? /p : *p."""
    out = np.array(logits, dtype=np.float32, copy=True)
    seen = set()
    for t in context:
        t = int(t)
        if t in seen:
            continue
        seen.add(t)
        if t < out.shape[0]:
            out[t] = out[t] / penalty if out[t] >= 0 else out[t] * penalty
    return out


def topk_order(values: np.ndarray, k: int) -> np.ndarray:
    """Total order of ops/fused_ops/portable.rs:28-32 / kernels/cuda/topk.cu: value
    descending, index ascending among equals."""
    v = np.asarray(values, dtype=np.float32)
    idx = np.lexsort((np.arange(v.shape[0]), -v.astype(np.float64)))
    return idx[:k].astype(np.uint32)


def build_causal_mask_rows(seq_len: int, start_pos: int):
    """qwen3/modeling.rs:1000-1014 mask predicate: key j visible to query i iff j <= start+i."""
    total = start_pos + seq_len
    return [[1 if j <= start_pos + i else 0 for j in range(total)] for i in range(seq_len)]
