"""CPU restatement of crane-serve's sampler -- TEST INFRASTRUCTURE ONLY.

Follows `sampling::sample` (crane-serve/src/engine/sampling.rs:169-380), `apply_penalties_inplace` (:422-480),
`sample_gumbel_max_idx` (:382-393) and the top-k total order of `crane_core::ops::topk_indices`
(crane-core/src/ops/fused_ops/portable.rs:28-66) on the GPU branch (`has_gpu_sampling`), in float32 like the candle tensors.
The uniforms of the Gumbel draw are an input (candle's device RNG stream is not pinned by any reference test).
Pinned by the reference's own literal cases: sampling.rs:489-640 (penalties), crane-core/tests/rocm_kernels.rs:141-198 (top-k).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def apply_penalties(logits, repetition_penalty=1.0, frequency_penalty=0.0, presence_penalty=0.0, context=()):
    """sampling.rs:422-480.  `x / p` and `x * p` on a candle tensor are affine ops: x * f32(1 / f64(p)) and x * f32(p)."""
    out = np.array(logits, dtype=f32, copy=True)
    rep = f32(repetition_penalty) != f32(1.0)
    fp = f32(frequency_penalty) != f32(0.0) or f32(presence_penalty) != f32(0.0)
    if len(context) == 0 or (not rep and not fp):
        return out
    counts = {}
    for t in context:
        counts[int(t)] = counts.get(int(t), 0) + 1
    inv = f32(1.0 / float(f32(repetition_penalty)))
    for t in sorted(counts):
        if t >= out.shape[0]:
            continue
        v = out[t]
        if rep:
            v = f32(v * inv) if v >= 0 else f32(v * f32(repetition_penalty))
        if fp:
            v = f32(v - f32(f32(counts[t]) * f32(frequency_penalty) + f32(presence_penalty)))
        out[t] = v
    return out


def topk_indices(values, k):
    """portable.rs:28-66: value descending, index ascending among equals."""
    v = np.asarray(values, dtype=f32)
    idx = np.lexsort((np.arange(v.shape[0]), -v.astype(np.float64)))
    return idx[:k].astype(np.uint32)


def gumbel_argmax(logits, temperature, uniforms):
    """sampling.rs:382-393: argmax(l / T - log(-log u)) (first maximum); `/ T` is an affine multiplication by f32(1 / T)."""
    l = np.asarray(logits, dtype=f32)
    if temperature <= 0:
        return int(np.flatnonzero(l == l.max())[0])
    u = np.asarray(uniforms, dtype=f32)[: l.shape[0]]
    minus_g = np.log(-np.log(u)).astype(f32)
    s = l if f32(temperature) == f32(1.0) else (l * f32(1.0 / float(f32(temperature)))).astype(f32)
    s = (s - minus_g).astype(f32)
    return int(np.flatnonzero(s == s.max())[0])


def sample(logits, temperature, top_p=None, top_k=None, repetition_penalty=1.0, frequency_penalty=0.0, presence_penalty=0.0,
           context=(), uniforms=None):
    """sampling.rs:169-380 on the GPU branch with CRANE_FORCE_GPU_TOPK=1 semantics for large vocabularies.
    Returns (token, logits after penalties)."""
    l = apply_penalties(logits, repetition_penalty, frequency_penalty, presence_penalty, context)
    vocab = l.shape[0]
    if temperature is None or temperature <= 0:
        return int(np.flatnonzero(l == l.max())[0]), l
    p = 1.0 if top_p is None else top_p
    top_p_active = 0.0 < p < 1.0
    k = top_k or 0
    if k == 0 and top_p_active:
        k = 64
    k = min(k, 64, vocab)
    if k > 0 and (k < vocab or top_p_active):
        idx = topk_indices(l, k)
        tl = l[idx]
        if top_p_active:
            scaled = (tl * f32(1.0 / float(f32(temperature)))).astype(f32)
            e = np.exp((scaled - scaled.max()).astype(f32)).astype(f32)
            s = f32(0)
            for x in e:
                s = f32(s + x)
            probs = (e / s).astype(f32)
            cum = np.zeros(k, f32)
            acc = f32(0)
            for j in range(k):
                acc = f32(acc + probs[j])
                cum[j] = acc
            le = cum <= f32(p)
            shift = np.zeros(k, bool)
            shift[1:] = le[:-1]
            masked = np.where(le | shift, tl, f32(-1e9)).astype(f32)
            pos = gumbel_argmax(masked, temperature, uniforms)
        else:
            pos = gumbel_argmax(tl, temperature, uniforms)
        return int(idx[pos]), l
    return gumbel_argmax(l, temperature, uniforms), l
