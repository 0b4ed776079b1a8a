"""ggml block formats Q8_0 / Q4_K / Q6_K: numpy dequantize + simple valid quantizers -- TEST INFRASTRUCTURE ONLY.

The reference keeps GGUF weights in these blocks and multiplies through candle's `QMatMul`
(crane-core/src/ops/linear.rs:23-48; hunyuan_dense/modeling.rs:37-41).  candle is not vendored, so the byte layouts are
restated from the published ggml definitions and PINNED against the `gguf` Python package (gguf.quants.dequantize /
quantize, installed here) in tests/test_oracle_golden.py.

The quantised LINEAR (`qmatmul` below) restates what candle's CPU `QMatMul::forward` computes behind
`LinearLayer::Quantized` (crane-core/src/ops/linear.rs:23-48): activations are quantised row by row to the weight type's
`VecDotType` -- Q8_K blocks (256 elements, iscale = -128 / max, q = min(127, round(iscale x)), d = 1 / iscale, per-16 sums) for
the K-quants, Q8_0 blocks (32 elements, d = amax / 127 stored as f16, q = round(x / d)) for Q8_0 weights -- and every output is
the published ggml integer dot `vec_dot_q4_K_q8_K` / `vec_dot_q6_K_q8_K` / `vec_dot_q8_0_q8_0`.  candle is a crates.io
dependency that is not vendored (Cargo.toml:12-14, "0.11"), so this is a restatement of the published ggml / candle k_quants
algorithms; it is pinned by identity checks (integer dot == dequant(W) . dequant(Q8(x)) in f64) and by the `gguf` package for
the weight side, NOT by a candle binary: "parity unpinned" for the activation rounding rule (round-half-away, as Rust's
f32::round) stays in force.

Layouts (little endian):
  Q8_0 : 32 elems / 34 B  : f16 d | 32 x i8                      y = d * q
  Q4_K : 256 elems / 144 B: f16 d | f16 dmin | 12 B packed 6-bit (scale, min) x 8 | 128 B nibbles
         sub-block j (32 elems): y = d*sc_j*q - dmin*m_j ; bytes [32p, 32p+32) hold sub-block 2p (low nibble) and 2p+1 (high)
  Q6_K : 256 elems / 210 B: 128 B ql | 64 B qh | 16 x i8 scales | f16 d      y = d * sc_(i/16) * (q - 32)
"""
from __future__ import annotations

import numpy as np

QK_K = 256
BLOCK_BYTES = {"Q8_0": 34, "Q4_K": 144, "Q6_K": 210}
BLOCK_ELEMS = {"Q8_0": 32, "Q4_K": 256, "Q6_K": 256}
GGML_TYPE_ID = {"Q8_0": 8, "Q4_K": 12, "Q6_K": 14}     # ggml_type enum values


def row_bytes(qtype: str, k: int) -> int:
    return k // BLOCK_ELEMS[qtype] * BLOCK_BYTES[qtype]


# ---------------------------------------------------------------------------------------------- Q8_0
def quantize_q8_0(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    n, k = x.shape
    b = x.reshape(n, k // 32, 32)
    amax = np.abs(b).max(-1)
    d = (amax / 127.0).astype(np.float16)
    df = d.astype(np.float32)
    inv = np.where(df > 0, 1.0 / np.where(df > 0, df, 1), 0).astype(np.float32)
    q = np.rint(b * inv[..., None]).clip(-127, 127).astype(np.int8)
    out = np.empty((n, k // 32, 34), np.uint8)
    out[..., :2] = d.view(np.uint8).reshape(n, k // 32, 2)
    out[..., 2:] = q.view(np.uint8)
    return out.reshape(n, -1)


def dequantize_q8_0(raw: np.ndarray, k: int) -> np.ndarray:
    n = raw.shape[0]
    b = raw.reshape(n, k // 32, 34)
    d = b[..., :2].copy().view(np.float16).astype(np.float32)[..., 0]
    q = b[..., 2:].view(np.int8).astype(np.float32)
    return (q * d[..., None]).reshape(n, k)


# ---------------------------------------------------------------------------------------------- Q4_K
def _pack_scales_k4(sc: np.ndarray, mn: np.ndarray) -> np.ndarray:
    """6-bit (scale, min) x 8 -> 12 bytes (ggml get_scale_min_k4 layout)."""
    out = np.zeros(sc.shape[:-1] + (12,), np.uint8)
    for j in range(4):
        out[..., j] = (sc[..., j] & 63) | ((sc[..., j + 4] >> 4) << 6)
        out[..., j + 4] = (mn[..., j] & 63) | ((mn[..., j + 4] >> 4) << 6)
        out[..., j + 8] = (sc[..., j + 4] & 0xF) | ((mn[..., j + 4] & 0xF) << 4)
    return out


def _unpack_scales_k4(s: np.ndarray):
    sc = np.zeros(s.shape[:-1] + (8,), np.uint8)
    mn = np.zeros_like(sc)
    for j in range(4):
        sc[..., j] = s[..., j] & 63
        mn[..., j] = s[..., j + 4] & 63
        sc[..., j + 4] = (s[..., j + 8] & 0xF) | ((s[..., j] >> 6) << 4)
        mn[..., j + 4] = (s[..., j + 8] >> 4) | ((s[..., j + 4] >> 6) << 4)
    return sc, mn


def quantize_q4_k(x: np.ndarray) -> np.ndarray:
    """A valid (not llama.cpp-optimal) Q4_K encoder: per sub-block affine 4-bit, 6-bit super-block scales."""
    x = np.ascontiguousarray(x, np.float32)
    n, k = x.shape
    b = x.reshape(n, k // 256, 8, 32)
    mn = np.maximum(0.0, -b.min(-1))                      # y = scale*q - min  with min >= 0
    scale = (b.max(-1) + mn) / 15.0
    d = (scale.max(-1) / 63.0).astype(np.float16)
    dmin = (mn.max(-1) / 63.0).astype(np.float16)
    df, dmf = d.astype(np.float32), dmin.astype(np.float32)
    sc = np.where(df[..., None] > 0, np.rint(scale / np.where(df[..., None] > 0, df[..., None], 1)), 0).clip(0, 63).astype(np.uint8)
    m6 = np.where(dmf[..., None] > 0, np.rint(mn / np.where(dmf[..., None] > 0, dmf[..., None], 1)), 0).clip(0, 63).astype(np.uint8)
    eff = df[..., None] * sc.astype(np.float32)
    q = np.where(eff[..., None] > 0, np.rint((b + (dmf[..., None] * m6)[..., None]) / np.where(eff[..., None] > 0, eff[..., None], 1)), 0)
    q = q.clip(0, 15).astype(np.uint8)                    # [n, nb, 8, 32]
    qs = (q[:, :, 0::2, :] | (q[:, :, 1::2, :] << 4)).reshape(n, k // 256, 128)
    out = np.empty((n, k // 256, 144), np.uint8)
    out[..., 0:2] = d.view(np.uint8).reshape(n, k // 256, 2)
    out[..., 2:4] = dmin.view(np.uint8).reshape(n, k // 256, 2)
    out[..., 4:16] = _pack_scales_k4(sc, m6)
    out[..., 16:] = qs
    return out.reshape(n, -1)


def dequantize_q4_k(raw: np.ndarray, k: int) -> np.ndarray:
    n = raw.shape[0]
    b = raw.reshape(n, k // 256, 144)
    d = b[..., 0:2].copy().view(np.float16).astype(np.float32)[..., 0]
    dmin = b[..., 2:4].copy().view(np.float16).astype(np.float32)[..., 0]
    sc, mn = _unpack_scales_k4(b[..., 4:16])
    qs = b[..., 16:].reshape(n, k // 256, 4, 32)
    q = np.empty((n, k // 256, 8, 32), np.float32)
    q[:, :, 0::2, :] = (qs & 0xF)
    q[:, :, 1::2, :] = (qs >> 4)
    y = (d[..., None] * sc.astype(np.float32))[..., None] * q - (dmin[..., None] * mn.astype(np.float32))[..., None]
    return y.reshape(n, k)


# ---------------------------------------------------------------------------------------------- Q6_K
def quantize_q6_k(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    n, k = x.shape
    nb = k // 256
    b = x.reshape(n, nb, 16, 16)
    amax = np.abs(b).max(-1)                               # per 16-elem group
    gscale = amax / 31.0
    d = (gscale.max(-1) / 127.0).astype(np.float16)
    df = d.astype(np.float32)
    sc = np.where(df[..., None] > 0, np.rint(gscale / np.where(df[..., None] > 0, df[..., None], 1)), 0).clip(-128, 127).astype(np.int8)
    eff = df[..., None] * sc.astype(np.float32)
    q = np.where(eff[..., None] != 0, np.rint(b / np.where(eff[..., None] != 0, eff[..., None], 1)), 0).clip(-32, 31).astype(np.int8)
    u = (q.reshape(n, nb, 256).astype(np.int16) + 32).astype(np.uint8)          # 0..63
    ql = np.zeros((n, nb, 128), np.uint8)
    qh = np.zeros((n, nb, 64), np.uint8)
    for half in range(2):
        e = u[:, :, 128 * half:128 * half + 128]
        l = np.arange(32)
        q1, q2, q3, q4 = e[:, :, l], e[:, :, l + 32], e[:, :, l + 64], e[:, :, l + 96]
        ql[:, :, 64 * half + l] = (q1 & 0xF) | ((q3 & 0xF) << 4)
        ql[:, :, 64 * half + l + 32] = (q2 & 0xF) | ((q4 & 0xF) << 4)
        qh[:, :, 32 * half + l] = (q1 >> 4) | ((q2 >> 4) << 2) | ((q3 >> 4) << 4) | ((q4 >> 4) << 6)
    out = np.empty((n, nb, 210), np.uint8)
    out[..., 0:128] = ql
    out[..., 128:192] = qh
    out[..., 192:208] = sc.view(np.uint8)
    out[..., 208:210] = d.view(np.uint8).reshape(n, nb, 2)
    return out.reshape(n, -1)


def dequantize_q6_k(raw: np.ndarray, k: int) -> np.ndarray:
    n = raw.shape[0]
    nb = k // 256
    b = raw.reshape(n, nb, 210)
    ql, qh = b[..., 0:128], b[..., 128:192]
    sc = b[..., 192:208].view(np.int8).astype(np.float32)
    d = b[..., 208:210].copy().view(np.float16).astype(np.float32)[..., 0]
    y = np.empty((n, nb, 256), np.float32)
    l = np.arange(32)
    for half in range(2):
        qlh, qhh = ql[:, :, 64 * half:64 * half + 64], qh[:, :, 32 * half:32 * half + 32]
        q1 = ((qlh[:, :, l] & 0xF) | (((qhh[:, :, l] >> 0) & 3) << 4)).astype(np.int16) - 32
        q2 = ((qlh[:, :, l + 32] & 0xF) | (((qhh[:, :, l] >> 2) & 3) << 4)).astype(np.int16) - 32
        q3 = ((qlh[:, :, l] >> 4) | (((qhh[:, :, l] >> 4) & 3) << 4)).astype(np.int16) - 32
        q4 = ((qlh[:, :, l + 32] >> 4) | (((qhh[:, :, l] >> 6) & 3) << 4)).astype(np.int16) - 32
        isb = 8 * half + l // 16
        base = 128 * half
        y[:, :, base + l] = d[..., None] * sc[:, :, isb + 0] * q1
        y[:, :, base + l + 32] = d[..., None] * sc[:, :, isb + 2] * q2
        y[:, :, base + l + 64] = d[..., None] * sc[:, :, isb + 4] * q3
        y[:, :, base + l + 96] = d[..., None] * sc[:, :, isb + 6] * q4
    return y.reshape(n, k)


QUANTIZE = {"Q8_0": quantize_q8_0, "Q4_K": quantize_q4_k, "Q6_K": quantize_q6_k}
DEQUANTIZE = {"Q8_0": dequantize_q8_0, "Q4_K": dequantize_q4_k, "Q6_K": dequantize_q6_k}


def quantize(x: np.ndarray, qtype: str) -> np.ndarray:
    return QUANTIZE[qtype](x)


def dequantize(raw: np.ndarray, qtype: str, k: int) -> np.ndarray:
    return DEQUANTIZE[qtype](np.ascontiguousarray(raw, np.uint8), k)


# ---------------------------------------------------------------------------------------------- activation blocks + integer dots
def _round_half_away(x: np.ndarray) -> np.ndarray:
    """Rust `f32::round` / C `roundf`: halves away from zero (np.rint is half-to-even)."""
    return np.sign(x) * np.floor(np.abs(x) + np.float32(0.5))


def quantize_act_q8_k(x: np.ndarray):
    """candle `BlockQ8K::from_float` / ggml `quantize_row_q8_K_ref`: x [m, k] f32 -> (q int8 [m, k/256, 256], d f32 [m, k/256],
    bsums int32 [m, k/256, 16]).  `max` is the FIRST element of largest magnitude, sign kept; iscale = -128 / max."""
    x = np.ascontiguousarray(x, np.float32)
    m, k = x.shape
    b = x.reshape(m, k // 256, 256)
    idx = np.abs(b).argmax(-1)                                  # first occurrence of the maximum magnitude
    mx = np.take_along_axis(b, idx[..., None], -1)[..., 0]
    nz = mx != 0
    iscale = np.where(nz, np.float32(-128.0) / np.where(nz, mx, 1).astype(np.float32), 0).astype(np.float32)
    v = _round_half_away((iscale[..., None] * b).astype(np.float32))
    q = np.minimum(v, 127).astype(np.int8)
    d = np.where(nz, np.float32(1.0) / np.where(nz, iscale, 1), 0).astype(np.float32)
    bsums = q.reshape(m, k // 256, 16, 16).astype(np.int32).sum(-1)
    return q, d, bsums


def quantize_act_q8_0(x: np.ndarray):
    """candle `BlockQ8_0::from_float` / ggml `quantize_row_q8_0_ref`: d = amax / 127 (q uses the f32 value, the block stores
    the f16 rounding of it), q = round(x * (1 / d)).  Returns (q int8 [m, k/32, 32], d_f16_as_f32 [m, k/32])."""
    x = np.ascontiguousarray(x, np.float32)
    m, k = x.shape
    b = x.reshape(m, k // 32, 32)
    d = (np.abs(b).max(-1) / np.float32(127.0)).astype(np.float32)
    inv = np.where(d != 0, np.float32(1.0) / np.where(d != 0, d, 1), 0).astype(np.float32)
    q = _round_half_away((b * inv[..., None]).astype(np.float32)).astype(np.int8)
    return q, d.astype(np.float16).astype(np.float32)


def dequantize_act(x: np.ndarray, qtype: str) -> np.ndarray:
    """What the quantised activations stand for: d * q, shaped like x (f32)."""
    if qtype == "Q8_0":
        q, d = quantize_act_q8_0(x)
    else:
        q, d, _ = quantize_act_q8_k(x)
    return (q.astype(np.float32) * d[..., None]).reshape(x.shape)


def qmatmul(x: np.ndarray, raw: np.ndarray, qtype: str) -> np.ndarray:
    """y[m, n] = ggml vec_dot(row n of the quantised weight, quantised row m of x): candle CPU `QMatMul::forward`.
    raw: [n, row_bytes] uint8 ggml blocks.  Integer sums are exact (int64); the float combination follows ggml's order
    (d = y.d * x.d, then d * isum, minus dmin * sum(mins * bsums))."""
    x = np.ascontiguousarray(x, np.float32)
    raw = np.ascontiguousarray(raw, np.uint8)
    m, k = x.shape
    n = raw.shape[0]
    if qtype == "Q8_0":
        qx, dx = quantize_act_q8_0(x)                                        # [m, nb, 32], [m, nb]
        b = raw.reshape(n, k // 32, 34)
        dw = b[..., :2].copy().view(np.float16).astype(np.float32)[..., 0]   # [n, nb]
        qw = b[..., 2:].view(np.int8)
        isum = np.einsum("mbk,nbk->mnb", qx.astype(np.int64), qw.astype(np.int64))
        return (isum.astype(np.float32) * (dx[:, None, :] * dw[None, :, :])).sum(-1, dtype=np.float32)
    qx, dx, bs = quantize_act_q8_k(x)                                        # [m, nb, 256], [m, nb], [m, nb, 16]
    nb = k // 256
    if qtype == "Q4_K":
        b = raw.reshape(n, nb, 144)
        dw = b[..., 0:2].copy().view(np.float16).astype(np.float32)[..., 0]
        dmin = b[..., 2:4].copy().view(np.float16).astype(np.float32)[..., 0]
        sc, mn = _unpack_scales_k4(b[..., 4:16])                             # [n, nb, 8] uint8
        qs = b[..., 16:].reshape(n, nb, 4, 32)
        q4 = np.empty((n, nb, 8, 32), np.int64)
        q4[:, :, 0::2, :] = qs & 0xF
        q4[:, :, 1::2, :] = qs >> 4
        sub = np.einsum("mbjk,nbjk->mnbj", qx.reshape(m, nb, 8, 32).astype(np.int64), q4)       # per 32-element sub-block
        isum = (sub * sc.astype(np.int64)[None]).sum(-1)                                        # [m, n, nb]
        bs32 = bs.reshape(m, nb, 8, 2).sum(-1).astype(np.int64)                                 # bsums[2j] + bsums[2j+1]
        summ = np.einsum("mbj,nbj->mnb", bs32, mn.astype(np.int64))
        d = dx[:, None, :] * dw[None, :, :]
        dm = dx[:, None, :] * dmin[None, :, :]
        return (d * isum.astype(np.float32) - dm * summ.astype(np.float32)).sum(-1, dtype=np.float32)
    if qtype == "Q6_K":
        b = raw.reshape(n, nb, 210)
        ql, qh = b[..., 0:128], b[..., 128:192]
        sc = b[..., 192:208].view(np.int8).astype(np.int64)                  # [n, nb, 16]
        dw = b[..., 208:210].copy().view(np.float16).astype(np.float32)[..., 0]
        q6 = np.empty((n, nb, 256), np.int64)
        l = np.arange(32)
        for half in range(2):
            qlh, qhh = ql[:, :, 64 * half:64 * half + 64], qh[:, :, 32 * half:32 * half + 32]
            base = 128 * half
            q6[:, :, base + l] = ((qlh[:, :, l] & 0xF) | (((qhh[:, :, l] >> 0) & 3) << 4)).astype(np.int64) - 32
            q6[:, :, base + l + 32] = ((qlh[:, :, l + 32] & 0xF) | (((qhh[:, :, l] >> 2) & 3) << 4)).astype(np.int64) - 32
            q6[:, :, base + l + 64] = ((qlh[:, :, l] >> 4) | (((qhh[:, :, l] >> 4) & 3) << 4)).astype(np.int64) - 32
            q6[:, :, base + l + 96] = ((qlh[:, :, l + 32] >> 4) | (((qhh[:, :, l] >> 6) & 3) << 4)).astype(np.int64) - 32
        sub = np.einsum("mbjk,nbjk->mnbj", qx.reshape(m, nb, 16, 16).astype(np.int64), q6.reshape(n, nb, 16, 16))
        isum = (sub * sc[None]).sum(-1)
        return ((dx[:, None, :] * dw[None, :, :]) * isum.astype(np.float32)).sum(-1, dtype=np.float32)
    raise ValueError(qtype)
