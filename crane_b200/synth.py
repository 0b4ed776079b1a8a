"""Model configurations (HF ``config.json`` form) and seeded synthetic checkpoints.

No checkpoints exist offline, so every test / bench run draws weights with the
rule the reference's own ``RandWeights`` test backend uses
(crane-core/src/models/qwen3_5/prefill.rs:162-173): per-tensor N(0, 1/fan_in),
rounded to bf16 -- the precision the safetensors hold and the precision the
reference's CPU path up-casts from (crane-core/src/models/qwen3/modeling.rs:1629-1632).

Tensor names are the HF safetensors names the reference loaders read
(crane-core/src/models/qwen3/modeling.rs:160-282,598-606,771-813;
 crane-core/src/models/qwen3_5/vision.rs:23-35,70-71,114-115,245-250,321-339).

numpy + zlib only: this module ships with the product (bench / tests feed the
C-ABI from it) and must not depend on the oracle.
"""
from __future__ import annotations

import zlib

import numpy as np

SEED = 20260923

# --------------------------------------------------------------------------------------
# Configurations.  Starred values in SURVEY.md section 8 are pinned by the reference,
# the rest are the public HF config.json values of each checkpoint.
# --------------------------------------------------------------------------------------

QWEN3_0_6B = {
    "model_type": "qwen3",
    "vocab_size": 151936, "hidden_size": 1024, "intermediate_size": 3072,
    "num_hidden_layers": 28, "num_attention_heads": 16, "num_key_value_heads": 8,
    "head_dim": 128, "max_position_embeddings": 40960, "rms_norm_eps": 1e-6,
    "rope_theta": 1000000.0, "tie_word_embeddings": True,
}

QWEN3_8B = {
    "model_type": "qwen3",
    "vocab_size": 151936, "hidden_size": 4096, "intermediate_size": 12288,
    "num_hidden_layers": 36, "num_attention_heads": 32, "num_key_value_heads": 8,
    "head_dim": 128, "max_position_embeddings": 40960, "rms_norm_eps": 1e-6,
    "rope_theta": 1000000.0, "tie_word_embeddings": False,
}

QWEN3_VL_2B = {
    "model_type": "qwen3_vl",
    "image_token_id": 151655, "vision_start_token_id": 151652, "vision_end_token_id": 151653,
    "tie_word_embeddings": True,
    "text_config": {
        "vocab_size": 151936, "hidden_size": 2048, "intermediate_size": 6144,
        "num_hidden_layers": 28, "num_attention_heads": 16, "num_key_value_heads": 8,
        "head_dim": 128, "max_position_embeddings": 262144, "rms_norm_eps": 1e-6,
        "rope_theta": 5000000.0, "tie_word_embeddings": True,
        "rope_scaling": {"mrope_section": [24, 20, 20], "mrope_interleaved": True},
    },
    "vision_config": {
        "depth": 24, "hidden_size": 1024, "intermediate_size": 4096, "num_heads": 16,
        "in_channels": 3, "patch_size": 16, "spatial_merge_size": 2, "temporal_patch_size": 2,
        "out_hidden_size": 2048, "num_position_embeddings": 2304,
        "deepstack_visual_indexes": [5, 11, 17], "hidden_act": "gelu_pytorch_tanh",
    },
}

# Tiny twins used by the parity tests (same structure, seconds on the CPU oracle).
TINY_QWEN3 = {
    "model_type": "qwen3",
    "vocab_size": 1024, "hidden_size": 256, "intermediate_size": 512,
    "num_hidden_layers": 3, "num_attention_heads": 4, "num_key_value_heads": 2,
    "head_dim": 128, "max_position_embeddings": 4096, "rms_norm_eps": 1e-6,
    "rope_theta": 1000000.0, "tie_word_embeddings": True,
}

TINY_QWEN3_UNTIED = dict(TINY_QWEN3, tie_word_embeddings=False, num_attention_heads=8,
                         num_key_value_heads=2, vocab_size=1000)

TINY_QWEN3_VL = {
    "model_type": "qwen3_vl",
    "image_token_id": 1001, "vision_start_token_id": 1002, "vision_end_token_id": 1003,
    "tie_word_embeddings": True,
    "text_config": {
        "vocab_size": 1024, "hidden_size": 256, "intermediate_size": 512,
        "num_hidden_layers": 4, "num_attention_heads": 4, "num_key_value_heads": 2,
        "head_dim": 128, "max_position_embeddings": 4096, "rms_norm_eps": 1e-6,
        "rope_theta": 5000000.0, "tie_word_embeddings": True,
        "rope_scaling": {"mrope_section": [24, 20, 20], "mrope_interleaved": True},
    },
    "vision_config": {
        "depth": 3, "hidden_size": 128, "intermediate_size": 256, "num_heads": 2,
        "in_channels": 3, "patch_size": 16, "spatial_merge_size": 2, "temporal_patch_size": 2,
        "out_hidden_size": 256, "num_position_embeddings": 64,
        "deepstack_visual_indexes": [0, 1, 2], "hidden_act": "gelu_pytorch_tanh",
    },
}


QWEN3_5_0_8B = {
    "model_type": "qwen3_5_text",
    "vocab_size": 248320, "hidden_size": 1024, "intermediate_size": 3584,
    "num_hidden_layers": 24, "num_attention_heads": 8, "num_key_value_heads": 2, "head_dim": 256,
    "max_position_embeddings": 262144, "rms_norm_eps": 1e-6, "tie_word_embeddings": True,
    "full_attention_interval": 4, "partial_rotary_factor": 0.25,
    "linear_conv_kernel_dim": 4, "linear_key_head_dim": 128, "linear_value_head_dim": 128,
    "linear_num_key_heads": 16, "linear_num_value_heads": 16,
    "rope_parameters": {"rope_type": "default", "rope_theta": 10000000.0, "partial_rotary_factor": 0.25,
                        "mrope_section": [11, 11, 10], "mrope_interleaved": True},
}

TINY_QWEN3_5 = {
    "model_type": "qwen3_5_text",
    "vocab_size": 1024, "hidden_size": 256, "intermediate_size": 512,
    "num_hidden_layers": 4, "num_attention_heads": 4, "num_key_value_heads": 2, "head_dim": 256,
    "max_position_embeddings": 4096, "rms_norm_eps": 1e-6, "tie_word_embeddings": True,
    "full_attention_interval": 4, "partial_rotary_factor": 0.25,
    "linear_conv_kernel_dim": 4, "linear_key_head_dim": 128, "linear_value_head_dim": 128,
    "linear_num_key_heads": 2, "linear_num_value_heads": 4,
    "rope_parameters": {"rope_type": "default", "rope_theta": 10000000.0, "partial_rotary_factor": 0.25,
                        "mrope_section": [11, 11, 10], "mrope_interleaved": True},
}


QWEN3_TTS_0_6B = {
    "model_type": "qwen3_tts",
    "tts_bos_token_id": 151672, "tts_eos_token_id": 151673, "tts_pad_token_id": 151671,
    "talker_config": {
        "vocab_size": 3072, "hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 28,
        "num_attention_heads": 16, "num_key_value_heads": 8, "head_dim": 128, "rms_norm_eps": 1e-6, "rope_theta": 1000000.0,
        "num_code_groups": 16, "text_hidden_size": 2048, "text_vocab_size": 151936, "max_position_embeddings": 32768,
        "codec_eos_token_id": 2150, "codec_think_id": 2154, "codec_nothink_id": 2155, "codec_think_bos_id": 2156,
        "codec_think_eos_id": 2157, "codec_pad_id": 2148, "codec_bos_id": 2149,
        "code_predictor_config": {
            "vocab_size": 2048, "hidden_size": 1024, "intermediate_size": 3072, "num_hidden_layers": 5,
            "num_attention_heads": 16, "num_key_value_heads": 8, "head_dim": 128, "rms_norm_eps": 1e-6, "rope_theta": 1000000.0,
            "num_code_groups": 16, "max_position_embeddings": 32768,
        },
    },
}

TINY_QWEN3_TTS = {
    "model_type": "qwen3_tts",
    "tts_bos_token_id": 1021, "tts_eos_token_id": 1022, "tts_pad_token_id": 1020,
    "talker_config": {
        "vocab_size": 1280, "hidden_size": 256, "intermediate_size": 512, "num_hidden_layers": 2,
        "num_attention_heads": 4, "num_key_value_heads": 2, "head_dim": 128, "rms_norm_eps": 1e-6, "rope_theta": 1000000.0,
        "num_code_groups": 4, "text_hidden_size": 512, "text_vocab_size": 1024, "max_position_embeddings": 4096,
        "codec_eos_token_id": 1270, "codec_think_id": 1274, "codec_nothink_id": 1275, "codec_think_bos_id": 1276,
        "codec_think_eos_id": 1277, "codec_pad_id": 1268, "codec_bos_id": 1269,
        "code_predictor_config": {
            "vocab_size": 256, "hidden_size": 256, "intermediate_size": 512, "num_hidden_layers": 2,
            "num_attention_heads": 4, "num_key_value_heads": 2, "head_dim": 128, "rms_norm_eps": 1e-6, "rope_theta": 1000000.0,
            "num_code_groups": 4, "max_position_embeddings": 4096,
        },
    },
}


def text_config(cfg: dict) -> dict:
    return cfg.get("text_config", cfg)


def head_dim(tc: dict) -> int:
    return tc.get("head_dim") or tc["hidden_size"] // tc["num_attention_heads"]


# --------------------------------------------------------------------------------------
# bf16 helpers (numpy has no bf16: carry it as uint16 bit patterns)
# --------------------------------------------------------------------------------------

def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    return (rounded >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """f32 values representable in bf16 (what the reference CPU path up-casts)."""
    return bf16_bits_to_f32(f32_to_bf16_bits(x))


# --------------------------------------------------------------------------------------
# Synthetic checkpoints
# --------------------------------------------------------------------------------------

def _rng(name: str) -> np.random.Generator:
    return np.random.default_rng(SEED + zlib.crc32(name.encode()))


def _normal(name: str, shape, std: float, mean: float = 0.0) -> np.ndarray:
    x = _rng(name).standard_normal(size=shape, dtype=np.float32)
    if std != 1.0:
        x *= np.float32(std)
    if mean != 0.0:
        x += np.float32(mean)
    return x


def text_tensor_specs(tc: dict, prefix: str = "model.", tie: bool | None = None):
    """Yield (name, shape, kind) for the dense Qwen3 decoder."""
    H, I = tc["hidden_size"], tc["intermediate_size"]
    nh, nkv, d, V = tc["num_attention_heads"], tc["num_key_value_heads"], head_dim(tc), tc["vocab_size"]
    yield prefix + "embed_tokens.weight", (V, H), "embed"
    for i in range(tc["num_hidden_layers"]):
        p = f"{prefix}layers.{i}."
        yield p + "input_layernorm.weight", (H,), "norm"
        yield p + "self_attn.q_proj.weight", (nh * d, H), "linear"
        yield p + "self_attn.k_proj.weight", (nkv * d, H), "linear"
        yield p + "self_attn.v_proj.weight", (nkv * d, H), "linear"
        yield p + "self_attn.q_norm.weight", (d,), "norm"
        yield p + "self_attn.k_norm.weight", (d,), "norm"
        yield p + "self_attn.o_proj.weight", (H, nh * d), "linear"
        yield p + "post_attention_layernorm.weight", (H,), "norm"
        yield p + "mlp.gate_proj.weight", (I, H), "linear"
        yield p + "mlp.up_proj.weight", (I, H), "linear"
        yield p + "mlp.down_proj.weight", (H, I), "linear"
    yield prefix + "norm.weight", (H,), "norm"
    tied = tc.get("tie_word_embeddings", True) if tie is None else tie
    if not tied:
        yield "lm_head.weight", (V, H), "linear"


def is_full_attention_layer(tc: dict, i: int) -> bool:
    """Layer i is softmax attention iff (i + 1) % full_attention_interval == 0 (qwen3_5/config.rs, HF layer_types)."""
    lt = tc.get("layer_types")
    if lt:
        return lt[i] == "full_attention"
    return (i + 1) % tc.get("full_attention_interval", 4) == 0


def qwen3_5_tensor_specs(tc: dict, prefix: str = "model."):
    """Hybrid Qwen3.5 text decoder: GDN (`linear_attn.*`) on 3 of 4 layers, gated softmax attention on the 4th.
    Names as read at crane-core/src/ops/gdn/layer.rs:55-67, projection.rs:75-83, models/qwen3_5/modeling.rs:337-358,587-669."""
    H, I, V = tc["hidden_size"], tc["intermediate_size"], tc["vocab_size"]
    nh, nkv, d = tc["num_attention_heads"], tc["num_key_value_heads"], head_dim(tc)
    nk, nv = tc["linear_num_key_heads"], tc["linear_num_value_heads"]
    dk, dv, ck = tc["linear_key_head_dim"], tc["linear_value_head_dim"], tc["linear_conv_kernel_dim"]
    key_dim, value_dim = nk * dk, nv * dv
    yield prefix + "embed_tokens.weight", (V, H), "embed"
    for i in range(tc["num_hidden_layers"]):
        p = f"{prefix}layers.{i}."
        yield p + "input_layernorm.weight", (H,), "norm0"
        if is_full_attention_layer(tc, i):
            yield p + "self_attn.q_proj.weight", (nh * d * 2, H), "linear"
            yield p + "self_attn.k_proj.weight", (nkv * d, H), "linear"
            yield p + "self_attn.v_proj.weight", (nkv * d, H), "linear"
            yield p + "self_attn.q_norm.weight", (d,), "norm0"
            yield p + "self_attn.k_norm.weight", (d,), "norm0"
            yield p + "self_attn.o_proj.weight", (H, nh * d), "linear"
        else:
            q = p + "linear_attn."
            yield q + "in_proj_qkv.weight", (2 * key_dim + value_dim, H), "linear"
            yield q + "in_proj_z.weight", (value_dim, H), "linear"
            yield q + "in_proj_b.weight", (nv, H), "linear"
            yield q + "in_proj_a.weight", (nv, H), "linear"
            yield q + "conv1d.weight", (2 * key_dim + value_dim, 1, ck), "linear"
            yield q + "dt_bias", (nv,), "dt_bias"
            yield q + "A_log", (nv,), "a_log"
            yield q + "norm.weight", (dv,), "norm"
            yield q + "out_proj.weight", (H, value_dim), "linear"
        yield p + "post_attention_layernorm.weight", (H,), "norm0"
        yield p + "mlp.gate_proj.weight", (I, H), "linear"
        yield p + "mlp.up_proj.weight", (I, H), "linear"
        yield p + "mlp.down_proj.weight", (H, I), "linear"
    yield prefix + "norm.weight", (H,), "norm0"
    if not tc.get("tie_word_embeddings", True):
        yield "lm_head.weight", (V, H), "linear"


def vision_tensor_specs(vc: dict, prefix: str = "model.visual."):
    Hv, Iv = vc["hidden_size"], vc["intermediate_size"]
    m2 = vc["spatial_merge_size"] ** 2
    P, T, C = vc["patch_size"], vc["temporal_patch_size"], vc["in_channels"]
    yield prefix + "patch_embed.proj.weight", (Hv, C, T, P, P), "linear"
    yield prefix + "patch_embed.proj.bias", (Hv,), "bias"
    yield prefix + "pos_embed.weight", (vc["num_position_embeddings"], Hv), "bias"
    for i in range(vc["depth"]):
        p = f"{prefix}blocks.{i}."
        for n in ("norm1", "norm2"):
            yield p + n + ".weight", (Hv,), "norm"
            yield p + n + ".bias", (Hv,), "bias"
        yield p + "attn.qkv.weight", (3 * Hv, Hv), "linear"
        yield p + "attn.qkv.bias", (3 * Hv,), "bias"
        yield p + "attn.proj.weight", (Hv, Hv), "linear"
        yield p + "attn.proj.bias", (Hv,), "bias"
        yield p + "mlp.linear_fc1.weight", (Iv, Hv), "linear"
        yield p + "mlp.linear_fc1.bias", (Iv,), "bias"
        yield p + "mlp.linear_fc2.weight", (Hv, Iv), "linear"
        yield p + "mlp.linear_fc2.bias", (Hv,), "bias"
    mergers = [("merger.", False)] + [
        (f"deepstack_merger_list.{j}.", True) for j in range(len(vc.get("deepstack_visual_indexes", [])))
    ]
    for name, post in mergers:
        p = prefix + name
        nd = Hv * m2 if post else Hv
        yield p + "norm.weight", (nd,), "norm"
        yield p + "norm.bias", (nd,), "bias"
        yield p + "linear_fc1.weight", (Hv * m2, Hv * m2), "linear"
        yield p + "linear_fc1.bias", (Hv * m2,), "bias"
        yield p + "linear_fc2.weight", (vc["out_hidden_size"], Hv * m2), "linear"
        yield p + "linear_fc2.bias", (vc["out_hidden_size"],), "bias"


def qwen3_tts_tensor_specs(cfg: dict):
    """Qwen3-TTS talker + code predictor (names: crane-core/src/models/qwen3_tts/modeling.rs:297-345,513-575):
    `talker.model.*` backbone, `talker.codec_head`, `talker.text_projection.*`, `talker.code_predictor.*`."""
    tk = cfg["talker_config"]
    cp = tk["code_predictor_config"]
    H, Ht = tk["hidden_size"], tk["text_hidden_size"]
    yield "talker.model.codec_embedding.weight", (tk["vocab_size"], H), "embed"
    yield "talker.model.text_embedding.weight", (tk["text_vocab_size"], Ht), "embed"
    yield "talker.text_projection.linear_fc1.weight", (Ht, Ht), "linear"
    yield "talker.text_projection.linear_fc1.bias", (Ht,), "bias"
    yield "talker.text_projection.linear_fc2.weight", (H, Ht), "linear"
    yield "talker.text_projection.linear_fc2.bias", (H,), "bias"
    for name, shape, kind in text_tensor_specs(dict(tk, tie_word_embeddings=True), "talker.model."):
        if "embed_tokens" not in name:
            yield name, shape, kind
    yield "talker.codec_head.weight", (tk["vocab_size"], H), "linear"
    n = cp["num_code_groups"] - 1
    for g in range(n):
        yield f"talker.code_predictor.model.codec_embedding.{g}.weight", (cp["vocab_size"], H), "embed"
    for name, shape, kind in text_tensor_specs(dict(cp, tie_word_embeddings=True), "talker.code_predictor.model."):
        if "embed_tokens" not in name:
            yield name, shape, kind
    for g in range(n):
        yield f"talker.code_predictor.lm_head.{g}.weight", (cp["vocab_size"], cp["hidden_size"]), "linear"
    if cp["hidden_size"] != H:
        yield "talker.code_predictor.small_to_mtp_projection.weight", (cp["hidden_size"], H), "linear"
        yield "talker.code_predictor.small_to_mtp_projection.bias", (cp["hidden_size"],), "bias"


def tensor_specs(cfg: dict):
    if cfg.get("model_type") == "qwen3_tts":
        yield from qwen3_tts_tensor_specs(cfg)
        return
    if cfg.get("model_type", "").startswith("qwen3_5"):
        yield from qwen3_5_tensor_specs(cfg)
    elif cfg.get("model_type") == "qwen3_vl":
        yield from text_tensor_specs(cfg["text_config"], "model.language_model.",
                                     tie=cfg.get("tie_word_embeddings", True))
        yield from vision_tensor_specs(cfg["vision_config"])
    else:
        yield from text_tensor_specs(cfg)


def make_tensor(name: str, shape, kind: str) -> np.ndarray:
    """One synthetic tensor as bf16-representable f32 values."""
    if kind == "linear":
        fan_in = int(np.prod(shape[1:]))
        x = _normal(name, shape, fan_in ** -0.5)
    elif kind == "embed":
        # unit-variance logits through the tied head after the final RMSNorm
        x = _normal(name, shape, shape[1] ** -0.5)
    elif kind == "norm":
        x = _normal(name, shape, 0.02, 1.0)
    elif kind == "norm0":      # Qwen3.5 stores w with (1 + w) applied at run time; real checkpoints sit near 0.24
        x = _normal(name, shape, 0.05, 0.24)
    elif kind == "a_log":      # keep the decay near 1 so state survives (qwen3_5/prefill.rs:164-171)
        x = _normal(name, shape, 0.1, -2.0)
    elif kind == "dt_bias":
        x = _normal(name, shape, 1.0)
    elif kind == "bias":
        x = _normal(name, shape, 0.02)
    else:
        raise ValueError(kind)
    return bf16_round(x)


def synth_checkpoint(cfg: dict):
    """Iterate (name, f32 ndarray with bf16-representable values)."""
    for name, shape, kind in tensor_specs(cfg):
        yield name, make_tensor(name, shape, kind)


def synth_token_ids(n: int, vocab: int, tag: str = "ids", forbid=()) -> np.ndarray:
    ids = _rng(tag).integers(0, vocab, size=n, dtype=np.int64)
    for f in forbid:
        ids[ids == f] = (f + 7) % vocab
    return ids.astype(np.uint32)


def synth_image(h: int, w: int, tag: str = "image") -> np.ndarray:
    """uint8 HWC image."""
    return _rng(tag).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


# --------------------------------------------------------------------------------------
# Host-side preprocessing that sits in front of the boundary (pure index arithmetic).
# --------------------------------------------------------------------------------------

def patchify(image_u8: np.ndarray, patch: int = 16, merge: int = 2, t_patch: int = 2,
             mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """uint8 HWC -> (pixel_values [n_patches, C*T*P*P] f32, grid_thw (1,h_p,w_p)).

    Row order is merge-block-major (h_blk, w_blk, m_row, m_col); each row is laid out
    (channel, temporal, patch_y, patch_x) with the still image duplicated along T.
    Follows crane-core/src/models/qwen3_5/processor.rs:114-209.  The image must already
    be at a smart_resize fixed point (sides multiples of patch*merge) -- resampling is
    host image processing and stays on the caller's side of the boundary.
    """
    h, w, c = image_u8.shape
    f = patch * merge
    if h % f or w % f:
        raise ValueError(f"image {h}x{w} is not a multiple of {f}; resize on the host first")
    chw = image_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    mean = np.asarray(mean, np.float32).reshape(3, 1, 1)
    std = np.asarray(std, np.float32).reshape(3, 1, 1)
    chw = (chw - mean) / std
    hp, wp = h // patch, w // patch
    x = chw.reshape(c, hp // merge, merge, patch, wp // merge, merge, patch)
    # -> (h_blk, w_blk, m_row, m_col, c, py, px)
    x = x.transpose(1, 4, 2, 5, 0, 3, 6)
    x = np.repeat(x[:, :, :, :, :, None, :, :], t_patch, axis=5)  # temporal duplicate
    pv = np.ascontiguousarray(x).reshape(hp * wp, c * t_patch * patch * patch)
    return pv.astype(np.float32), (1, hp, wp)


def build_vl_prompt(cfg: dict, n_text: int, grid_thw, tag: str = "vlprompt") -> np.ndarray:
    """<|vision_start|> + image_pad * (t*h/2*w/2) + <|vision_end|> + n_text random ids.

    Mirrors the placeholder expansion of crane-core/src/models/qwen3_5/vlm.rs:313-346
    (chat-template tokens other than the vision markers are ordinary text ids here).
    """
    tc = cfg["text_config"]
    m = cfg["vision_config"]["spatial_merge_size"]
    t, h, w = grid_thw
    n_img = t * (h // m) * (w // m)
    special = (cfg["image_token_id"], cfg["vision_start_token_id"], cfg["vision_end_token_id"])
    lo = min(special)
    text = synth_token_ids(n_text, min(tc["vocab_size"], lo), tag)
    n_pre = n_text // 8
    ids = np.concatenate([
        text[:n_pre],
        np.array([cfg["vision_start_token_id"]], np.uint32),
        np.full(n_img, cfg["image_token_id"], np.uint32),
        np.array([cfg["vision_end_token_id"]], np.uint32),
        text[n_pre:],
    ])
    return ids.astype(np.uint32)
