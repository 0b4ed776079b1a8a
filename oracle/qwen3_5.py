"""CPU oracle for the hybrid Qwen3.5 text decoder (Gated-Delta-Net + gated softmax attention)
-- TEST INFRASTRUCTURE ONLY (see oracle/qwen3.py for the rules and the pinning story).

Reference arithmetic restated (all under crane-core/src/):
  Qwen35RmsNorm (w <- 1 + w) ..................... models/qwen3_5/modeling.rs:45-79
  MRotaryEmbedding (f32 inv_freq, partial rotary)  models/qwen3_5/modeling.rs:98-279
  FullAttention (q_proj = per-head [query|gate],
     y *= sigmoid(gate), GQA) .................... models/qwen3_5/modeling.rs:413-564
  Mlp (3 separate linears) ....................... models/qwen3_5/modeling.rs:622-628
  DecoderLayer / Qwen3_5TextModel ................ models/qwen3_5/modeling.rs:784-832, model.rs:395-510
  chunked prefill (chunk == single pass) ......... models/qwen3_5/prefill.rs:56-99
  GatedDeltaNet::forward ......................... ops/gdn/layer.rs:122-238
  causal_conv1d / decode_conv1d .................. ops/gdn/conv.rs:23-133
  l2_norm, softplus, compute_beta_g, recurrence .. ops/gdn/backend.rs:26-215  (kernel: kernels/cuda/gdn.cu:29-34)
  RmsNormGated ................................... ops/gdn/norm.rs:39-45
Second opinion: HF `Qwen3_5ForCausalLM` (torch_chunk / torch_recurrent gated delta rule) via oracle/make_golden.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from crane_b200.synth import head_dim as _head_dim, is_full_attention_layer
from .qwen3 import causal_attention, silu
from .qwen3_vl import mrope_cos_sin


def rms_norm_1p(x, w, eps):
    """Qwen35RmsNorm: x * rsqrt(mean(x^2)+eps) * (1 + w)."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + w)


def l2_norm(x, eps=1e-6):
    """ops/gdn/backend.rs:26-37: x / sqrt(sum(x^2) + eps)."""
    return x / torch.sqrt(x.pow(2).sum(-1, keepdim=True) + eps)


def softplus(x):
    """ops/gdn/backend.rs:66-68: log(1 + exp(x)), no threshold."""
    return torch.log(1.0 + torch.exp(x))


def rope_tables_f32(rot_dim: int, max_pos: int, theta: float):
    """MRotaryEmbedding::new (qwen3_5/modeling.rs:110-132): inv_freq computed in F32
    (`1 / base.powf(i * 2 / rot_dim)`), unlike the f64 recipe of the dense Qwen3 tables."""
    base = np.float32(theta)
    inv = np.array([np.float32(1.0) / np.power(base, np.float32(i) * np.float32(2.0) / np.float32(rot_dim), dtype=np.float32)
                    for i in range(rot_dim // 2)], dtype=np.float32)
    pos = np.arange(max_pos, dtype=np.float32)
    freqs = (pos[:, None] * inv[None, :]).astype(np.float32)
    return torch.from_numpy(np.cos(freqs)), torch.from_numpy(np.sin(freqs))


def apply_partial_rope(x, cos, sin, rot_dim):
    """apply_mrope (qwen3_5/modeling.rs:263-279): rotate the first rot_dim components (half-split inside the slice)."""
    xr, xp = x[..., :rot_dim], x[..., rot_dim:]
    h = rot_dim // 2
    x1, x2 = xr[..., :h], xr[..., h:]
    c, s = cos[:, None, :], sin[:, None, :]
    return torch.cat([x1 * c - x2 * s, x1 * s + x2 * c, xp], dim=-1)


def gated_delta_rule(q, k, v, g, beta, state):
    """gated_delta_rule_recurrence (ops/gdn/backend.rs:90-156), f32.  q,k [S,Hv,K] (already L2-normed; q is scaled
    by 1/sqrt(K) here), v [S,Hv,V], g, beta [S,Hv], state [Hv,K,V] -> (y [S,Hv,V], new state)."""
    S = q.shape[0]
    q = q * (1.0 / math.sqrt(q.shape[-1]))
    s = state.clone()
    ys = []
    for t in range(S):
        s = s * torch.exp(g[t])[:, None, None]
        kv = (s * k[t][:, :, None]).sum(1)                       # [Hv, V]
        delta = (v[t] - kv) * beta[t][:, None]
        s = s + k[t][:, :, None] * delta[:, None, :]
        ys.append((s * q[t][:, :, None]).sum(1))
    return torch.stack(ys, 0), s


class Qwen3_5Oracle:
    def __init__(self, cfg: dict, weights: dict, prefix="model.", max_pos=8192, kv_bits: int = 0):
        self.kv_bits = kv_bits          # 8 / 4: QuantKvCache on the full-attention layers (qwen3_5/kv_cache.rs:209-342)
        tc = cfg.get("text_config", cfg)
        self.tc = tc
        self.H, self.L, self.V = tc["hidden_size"], tc["num_hidden_layers"], tc["vocab_size"]
        self.nh, self.nkv, self.d = tc["num_attention_heads"], tc["num_key_value_heads"], _head_dim(tc)
        self.nk, self.nv = tc["linear_num_key_heads"], tc["linear_num_value_heads"]
        self.dk, self.dv, self.ck = tc["linear_key_head_dim"], tc["linear_value_head_dim"], tc["linear_conv_kernel_dim"]
        self.key_dim, self.value_dim = self.nk * self.dk, self.nv * self.dv
        self.conv_dim = 2 * self.key_dim + self.value_dim
        self.eps = float(tc.get("rms_norm_eps", 1e-6))
        rp = tc.get("rope_parameters", {})
        self.theta = float(rp.get("rope_theta", tc.get("rope_theta", 1e7)))
        self.rot_dim = int(self.d * float(rp.get("partial_rotary_factor", tc.get("partial_rotary_factor", 0.25))))
        self.section = rp.get("mrope_section", [])
        self.p = prefix
        self.w = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).float() for k, v in weights.items()}
        self.lm_head = self.w.get("lm_head.weight", self.w[prefix + "embed_tokens.weight"]) \
            if not tc.get("tie_word_embeddings", True) else self.w[prefix + "embed_tokens.weight"]
        self.cos, self.sin = rope_tables_f32(self.rot_dim, max_pos, self.theta)
        self.full = [is_full_attention_layer(tc, i) for i in range(self.L)]
        self.clear_kv_cache()

    def clear_kv_cache(self):
        self.k_cache = [None] * self.L
        self.v_cache = [None] * self.L
        # GdnLayerCache (ops/gdn/cache.rs:15-45): conv_state [conv_dim, kernel] zeros, recurrent state [Hv, K, V] zeros
        self.conv_state = [torch.zeros(self.conv_dim, self.ck) for _ in range(self.L)]
        self.rec_state = [torch.zeros(self.nv, self.dk, self.dv) for _ in range(self.L)]
        self.n_cached = 0

    def W(self, i, n):
        return self.w[f"{self.p}layers.{i}.{n}"]

    def embed(self, ids):
        return self.w[self.p + "embed_tokens.weight"][torch.as_tensor(np.asarray(ids, dtype=np.int64))]

    def _full_attention(self, i, h, cos, sin, kv_offset):
        S = h.shape[0]
        qg = (h @ self.W(i, "self_attn.q_proj.weight").T).view(S, self.nh, 2 * self.d)
        q, gate = qg[..., :self.d], qg[..., self.d:].reshape(S, self.nh * self.d)
        k = (h @ self.W(i, "self_attn.k_proj.weight").T).view(S, self.nkv, self.d)
        v = (h @ self.W(i, "self_attn.v_proj.weight").T).view(S, self.nkv, self.d)
        q = rms_norm_1p(q, self.W(i, "self_attn.q_norm.weight"), self.eps)
        k = rms_norm_1p(k, self.W(i, "self_attn.k_norm.weight"), self.eps)
        q = apply_partial_rope(q, cos, sin, self.rot_dim)
        k = apply_partial_rope(k, cos, sin, self.rot_dim)
        if self.kv_bits:
            from .qwen3 import quant_kv_per_token
            k, v = quant_kv_per_token(k, self.kv_bits), quant_kv_per_token(v, self.kv_bits)
        if self.k_cache[i] is None:
            self.k_cache[i], self.v_cache[i] = k, v
        else:
            self.k_cache[i] = torch.cat([self.k_cache[i], k], 0)
            self.v_cache[i] = torch.cat([self.v_cache[i], v], 0)
        a = causal_attention(q, self.k_cache[i], self.v_cache[i], kv_offset, 1.0 / math.sqrt(self.d))
        a = a * torch.sigmoid(gate)
        return a @ self.W(i, "self_attn.o_proj.weight").T

    def _gdn(self, i, h):
        S = h.shape[0]
        mixed = h @ self.W(i, "linear_attn.in_proj_qkv.weight").T            # [S, conv_dim]
        z = (h @ self.W(i, "linear_attn.in_proj_z.weight").T).view(S, self.nv, self.dv)
        b = h @ self.W(i, "linear_attn.in_proj_b.weight").T                  # [S, nv]
        a = h @ self.W(i, "linear_attn.in_proj_a.weight").T
        # causal depthwise conv over [cached tail | this call], then SiLU (ops/gdn/conv.rs:23-101)
        wconv = self.W(i, "linear_attn.conv1d.weight").reshape(self.conv_dim, self.ck)
        hidden = torch.cat([self.conv_state[i], mixed.T], 1)                  # [conv_dim, ck + S]
        self.conv_state[i] = hidden[:, -self.ck:].clone()
        win = hidden.unfold(1, self.ck, 1)[:, 1:, :]                          # window for t = hidden[:, 1+t : 1+t+ck]
        conv = silu((win * wconv[:, None, :]).sum(-1)).T                      # [S, conv_dim]
        q = conv[:, :self.key_dim].reshape(S, self.nk, self.dk)
        k = conv[:, self.key_dim:2 * self.key_dim].reshape(S, self.nk, self.dk)
        v = conv[:, 2 * self.key_dim:].reshape(S, self.nv, self.dv)
        rep = self.nv // self.nk
        if rep > 1:   # VHeadOrder::Interleaved (HF): value head = key_head * rep + r  (ops/gdn/layer.rs:194-238)
            q = q.repeat_interleave(rep, dim=1)
            k = k.repeat_interleave(rep, dim=1)
        q, k = l2_norm(q), l2_norm(k)
        beta = torch.sigmoid(b)
        g = -torch.exp(self.W(i, "linear_attn.A_log")) * softplus(a + self.W(i, "linear_attn.dt_bias"))
        y, self.rec_state[i] = gated_delta_rule(q, k, v, g, beta, self.rec_state[i])
        # RmsNormGated (ops/gdn/norm.rs:39-45): rmsnorm(y; w) * silu(z), plain weight
        yn = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + self.eps) * self.W(i, "linear_attn.norm.weight")
        o = (yn * silu(z)).reshape(S, self.value_dim)
        return o @ self.W(i, "linear_attn.out_proj.weight").T

    def forward_embeds(self, x, start_pos, pos3=None):
        S = x.shape[0]
        assert start_pos == self.n_cached
        if pos3 is None:
            pos3 = np.tile(np.arange(start_pos, start_pos + S, dtype=np.uint32), (3, 1))
        cos, sin = mrope_cos_sin(self.cos, self.sin, pos3, self.section)
        for i in range(self.L):
            h = rms_norm_1p(x, self.W(i, "input_layernorm.weight"), self.eps)
            x = x + (self._full_attention(i, h, cos, sin, start_pos) if self.full[i] else self._gdn(i, h))
            h = rms_norm_1p(x, self.W(i, "post_attention_layernorm.weight"), self.eps)
            g = h @ self.W(i, "mlp.gate_proj.weight").T
            u = h @ self.W(i, "mlp.up_proj.weight").T
            x = x + (silu(g) * u) @ self.W(i, "mlp.down_proj.weight").T
        self.n_cached += S
        x = rms_norm_1p(x, self.w[self.p + "norm.weight"], self.eps)
        return x[-1] @ self.lm_head.T

    def forward(self, ids, start_pos):
        return self.forward_embeds(self.embed(ids), start_pos)
