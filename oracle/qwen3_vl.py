"""CPU oracle for the Qwen3-VL prefix (ViT + splice + interleaved MRoPE + DeepStack)
-- TEST INFRASTRUCTURE ONLY (see oracle/qwen3.py for the rules and the pinning story).

The reference's `qwen3_vl` module is dead code (crane-core/src/models/mod.rs:19); per
SURVEY.md F4 the arithmetic to follow is (all crane-core/src/models/):
  ViT .............................. qwen3_5/vision.rs (whole file; identical to qwen3_vl/vision.rs)
  text decoder ..................... qwen3/modeling.rs (oracle/qwen3.py)
  interleaved MRoPE cos/sin ........ qwen3_5/modeling.rs:172-245
  3-axis position ids .............. qwen3_5/vlm.rs:190-241
  image-feature splice ............. qwen3_5/vlm.rs:433-468
  prefill / decode_step ............ qwen3_5/vlm.rs:250-301
  DeepStack injection .............. qwen3_vl/text.rs:252-270,280-333
Second opinion: HF `Qwen3VLForConditionalGeneration` (oracle/make_golden.py).

Activation note: the reference maps the ViT MLP's `gelu_pytorch_tanh` to
candle `Activation::Gelu`, which is the ERF gelu (qwen3_5/config.rs:160-172), and the patch
mergers call `xs.gelu()`, candle's TANH approximation (qwen3_5/vision.rs:276).  HF does the
opposite.  `vit_act` / `merger_act` default to the reference's choice; the HF cross-check
flips both.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .qwen3 import Qwen3Oracle, rope_tables


def layer_norm(x, w, b, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def gelu(x, kind):
    return F.gelu(x, approximate="tanh") if kind == "tanh" else F.gelu(x)


class VisionOracle:
    def __init__(self, vc: dict, weights: dict, prefix="model.visual.", vit_act="erf", merger_act="tanh"):
        self.vc = vc
        self.p = prefix
        self.w = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).float()
                  for k, v in weights.items() if k.startswith(prefix)}
        self.Hv = vc["hidden_size"]
        self.nh = vc["num_heads"]
        self.hd = self.Hv // self.nh
        self.merge = vc["spatial_merge_size"]
        self.grid_side = int(round(math.sqrt(vc["num_position_embeddings"])))
        self.ds_idx = list(vc.get("deepstack_visual_indexes", []))
        self.vit_act, self.merger_act = vit_act, merger_act

    def W(self, n):
        return self.w[self.p + n]

    # qwen3_5/vision.rs:382-489
    def pos_embed_interpolate(self, grid_thw):
        side = self.grid_side
        table = self.W("pos_embed.weight")
        outs = []
        for (t, h, w) in grid_thw:
            def lin(steps):
                if steps == 1:
                    return np.zeros(1, np.float32)
                step = np.float32(side - 1) / np.float32(steps - 1)
                return (np.arange(steps, dtype=np.float32) * step).astype(np.float32)
            hv, wv = lin(h), lin(w)
            hf, wf = np.floor(hv).astype(np.int64), np.floor(wv).astype(np.int64)
            hc = np.minimum(np.ceil(hv).astype(np.int64), side - 1)
            wc = np.minimum(np.ceil(wv).astype(np.int64), side - 1)
            dh, dw = (hv - hf.astype(np.float32)), (wv - wf.astype(np.float32))
            idx = [hf[:, None] * side + wf[None, :], hf[:, None] * side + wc[None, :],
                   hc[:, None] * side + wf[None, :], hc[:, None] * side + wc[None, :]]
            wts = [(1 - dh)[:, None] * (1 - dw)[None, :], (1 - dh)[:, None] * dw[None, :],
                   dh[:, None] * (1 - dw)[None, :], dh[:, None] * dw[None, :]]
            pe = torch.zeros(h * w, self.Hv)
            for ii, ww in zip(idx, wts):
                # weights are cast to the embedding dtype (bf16 on the GPU path; f32 here)
                pe = pe + table[torch.from_numpy(ii.reshape(-1))] * torch.from_numpy(
                    ww.astype(np.float32).reshape(-1, 1))
            m = self.merge
            pe = pe.repeat(t, 1).view(t, h // m, m, w // m, m, self.Hv).permute(0, 1, 3, 2, 4, 5)
            outs.append(pe.reshape(t * h * w, self.Hv))
        return torch.cat(outs, 0)

    # qwen3_5/vision.rs:281-304,491-541
    def rot_pos_emb(self, grid_thw):
        dim = self.hd // 2
        inv = np.array([np.float32(1.0) / np.float32(10000.0) ** (np.float32(i) / np.float32(dim))
                        for i in range(0, dim, 2)], dtype=np.float32)
        max_hw = max(max(h, w) for (_, h, w) in grid_thw)
        table = (np.arange(max_hw, dtype=np.float32)[:, None] * inv[None, :]).astype(np.float32)
        rows, cols = [], []
        m = self.merge
        for (t, h, w) in grid_thw:
            r, c = [], []
            for br in range(h // m):
                for bc in range(w // m):
                    for ir in range(m):
                        for ic in range(m):
                            r.append(br * m + ir)
                            c.append(bc * m + ic)
            rows += r * t
            cols += c * t
        emb = np.concatenate([table[rows], table[cols]], axis=-1)   # [N, hd/2]
        return torch.from_numpy(emb)

    def _attn(self, i, x, cu, cos, sin):
        N = x.shape[0]
        qkv = x @ self.W(f"blocks.{i}.attn.qkv.weight").T + self.W(f"blocks.{i}.attn.qkv.bias")
        qkv = qkv.view(N, 3, self.nh, self.hd)
        q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]

        def rot(t):  # q*cos + rotate_half(q)*sin with full-width cos/sin (vision.rs:83-102)
            h2 = self.hd // 2
            rh = torch.cat([-t[..., h2:], t[..., :h2]], -1)
            return t * cos[:, None, :] + rh * sin[:, None, :]
        q, k = rot(q), rot(k)
        outs = []
        for a, b in zip(cu[:-1], cu[1:]):
            qc, kc, vc = (t[a:b].permute(1, 0, 2) for t in (q, k, v))
            sc = torch.matmul(qc, kc.transpose(1, 2)) / math.sqrt(self.hd)
            o = torch.matmul(torch.softmax(sc, -1), vc)          # non-causal, per image
            outs.append(o.permute(1, 0, 2).reshape(b - a, self.Hv))
        o = torch.cat(outs, 0)
        return o @ self.W(f"blocks.{i}.attn.proj.weight").T + self.W(f"blocks.{i}.attn.proj.bias")

    def _merger(self, name, x, post):
        m2 = self.merge ** 2
        g = x.shape[0] // m2
        if post:
            x = x.reshape(g, self.Hv * m2)
        x = layer_norm(x, self.W(name + "norm.weight"), self.W(name + "norm.bias"))
        x = x.reshape(g, self.Hv * m2)
        x = x @ self.W(name + "linear_fc1.weight").T + self.W(name + "linear_fc1.bias")
        x = gelu(x, self.merger_act)
        return x @ self.W(name + "linear_fc2.weight").T + self.W(name + "linear_fc2.bias")

    # qwen3_5/vision.rs:558-584
    def forward(self, pixel_values, grid_thw):
        pv = torch.as_tensor(pixel_values).float()
        wpe = self.W("patch_embed.proj.weight").reshape(self.Hv, -1)
        x = pv @ wpe.T + self.W("patch_embed.proj.bias")          # Conv3d(k=stride) == GEMM
        x = x + self.pos_embed_interpolate(grid_thw)
        rp = self.rot_pos_emb(grid_thw)
        emb = torch.cat([rp, rp], -1)
        cos, sin = emb.cos(), emb.sin()
        cu = [0]
        for (t, h, w) in grid_thw:
            for _ in range(t):
                cu.append(cu[-1] + h * w)
        deep = []
        for i in range(self.vc["depth"]):
            h1 = layer_norm(x, self.W(f"blocks.{i}.norm1.weight"), self.W(f"blocks.{i}.norm1.bias"))
            x = x + self._attn(i, h1, cu, cos, sin)
            h2 = layer_norm(x, self.W(f"blocks.{i}.norm2.weight"), self.W(f"blocks.{i}.norm2.bias"))
            y = h2 @ self.W(f"blocks.{i}.mlp.linear_fc1.weight").T + self.W(f"blocks.{i}.mlp.linear_fc1.bias")
            y = gelu(y, self.vit_act)
            x = x + y @ self.W(f"blocks.{i}.mlp.linear_fc2.weight").T + self.W(f"blocks.{i}.mlp.linear_fc2.bias")
            if i in self.ds_idx:
                j = self.ds_idx.index(i)
                deep.append(self._merger(f"deepstack_merger_list.{j}.", x, True))
        return self._merger("merger.", x, False), deep


def build_position_ids(ids, grid_thw, merge, image_token_id, start_pos=0):
    """qwen3_5/vlm.rs:190-241 -> ([3, S] uint32, next_mrope_pos)."""
    ids = [int(t) for t in ids]
    S = len(ids)
    pos = np.zeros((3, S), dtype=np.uint32)
    nxt, img, i = start_pos, 0, 0
    while i < S:
        if ids[i] != image_token_id:
            pos[:, i] = nxt
            nxt += 1
            i += 1
            continue
        gt, gh, gw = grid_thw[img][0], grid_thw[img][1] // merge, grid_thw[img][2] // merge
        span, hw, base = gt * gh * gw, gh * gw, nxt
        if i + span > S:
            raise ValueError("image span exceeds the sequence")
        for k in range(span):
            pos[:, i + k] = (base + k // hw, base + (k % hw) // gw, base + (k % hw) % gw)
        nxt = base + max(gt, gh, gw)
        i += span
        img += 1
    return pos, nxt


def mrope_axis_of(half: int, section):
    """Column ownership of the interleaved MRoPE (qwen3_5/modeling.rs:203-233)."""
    axis = [0] * half
    for dim in (1, 2):
        sec = section[dim] if dim < len(section) else 0
        for i in range(dim, min(sec * 3, half), 3):
            axis[i] = dim
    return axis


def mrope_cos_sin(cos_t, sin_t, pos3, section):
    """cos_sin_with_position_ids (qwen3_5/modeling.rs:172-245): column i comes from axis_of[i]."""
    half = cos_t.shape[1]
    axis = torch.tensor(mrope_axis_of(half, section))
    p = torch.from_numpy(pos3.astype(np.int64))                 # [3, S]
    sel = p[axis, :].T                                          # [S, half] position per column
    col = torch.arange(half)[None, :].expand_as(sel)
    return cos_t[sel, col], sin_t[sel, col]


class Qwen3VLOracle:
    """Prefill with one or more images, then decode_step -- `Qwen3_5VLModel::{forward,decode_step}`
    (qwen3_5/vlm.rs:250-301) over the dense Qwen3 decoder with DeepStack (qwen3_vl/text.rs:252-270)."""

    def __init__(self, cfg, weights, vit_act="erf", merger_act="tanh", max_pos=8192):
        self.cfg = cfg
        tc, vc = cfg["text_config"], cfg["vision_config"]
        self.text = Qwen3Oracle(cfg, {k: v for k, v in weights.items() if not k.startswith("model.visual.")},
                                prefix="model.language_model.", max_pos=max_pos)
        self.vision = VisionOracle(vc, weights, vit_act=vit_act, merger_act=merger_act)
        self.section = tc.get("rope_scaling", {}).get("mrope_section", [])
        self.image_token_id = cfg["image_token_id"]
        self.merge = vc["spatial_merge_size"]
        self.next_mrope_pos = 0

    def clear_kv_cache(self):
        self.text.clear_kv_cache()
        self.next_mrope_pos = 0

    def prefill(self, ids, pixel_values=None, grid_thw=None, start_pos=0):
        ids = np.asarray(ids, dtype=np.int64)
        x = self.text.embed(ids)
        deep, vis_rows = [], None
        if pixel_values is not None:
            img, deep = self.vision.forward(pixel_values, grid_thw)
            vis_rows = np.nonzero(ids == self.image_token_id)[0]
            assert len(vis_rows) == img.shape[0], "placeholder / image-embedding count mismatch"
            x = x.clone()
            x[vis_rows] = img                                  # splice (vlm.rs:433-468)
            pos3, nxt = build_position_ids(ids, grid_thw, self.merge, self.image_token_id, start_pos)
        else:
            pos3 = np.tile(np.arange(start_pos, start_pos + len(ids), dtype=np.uint32), (3, 1))
            nxt = start_pos + len(ids)
        self.next_mrope_pos = nxt
        cs = mrope_cos_sin(self.text.cos, self.text.sin, pos3, self.section)

        def after(i, h):                                       # DeepStack (text.rs:262-268,280-333)
            if vis_rows is not None and i < len(deep):
                h = h.clone()
                h[vis_rows] = h[vis_rows] + deep[i]
            return h
        return self.text.forward_embeds(x, start_pos, cos_sin=cs, after_layer=after)

    def decode_step(self, token: int, start_pos: int):
        p = self.next_mrope_pos
        self.next_mrope_pos = p + 1
        pos3 = np.full((3, 1), p, dtype=np.uint32)
        cs = mrope_cos_sin(self.text.cos, self.text.sin, pos3, self.section)
        return self.text.forward_embeds(self.text.embed([token]), start_pos, cos_sin=cs)
