"""TEST INFRASTRUCTURE (oracle): the chunkwise (WY / UT-transform) form of the gated delta rule, restated in numpy.

The reference runs the recurrence token by token (`gated_delta_rule_recurrence`, crane-core/src/ops/gdn/backend.rs:90-156,
kernels/cuda/gdn.cu:29-34; sequential restatement: oracle/qwen3_5.py `gated_delta_rule`).  The CUDA prefill path
(`crane_b200/csrc/gdn_chunk.cu`) evaluates the same recurrence 64 tokens at a time; this file is the algebra it follows,
block for block, so that tests can check (a) the algebra against the sequential rule and (b) every intermediate of the
kernels (W, U, K~, Q~, P, per-chunk start states, D) against a plain f64 / f32 statement.

Per value head, with local index i = 0..C-1 inside a chunk, start state S0 [K, V], a_i = exp(g_i), b_i = sum_{j<=i} g_j:
    S_i = a_i S_{i-1} + k_i d_i^T,     d_i = beta_i (v_i - (a_i S_{i-1})^T k_i),     y_i = S_i^T q_i
 => (I + A) D = diag(beta) V - diag(beta e^b) K S0,   A_ij = beta_i e^{b_i - b_j} (k_i . k_j)  for j < i
    W = (I + A)^-1 diag(beta e^b) K,  U = (I + A)^-1 diag(beta) V,  D = U - W S0
    Y = diag(e^b) Q S0 + P D,          P_ij = e^{b_i - b_j} (q_i . k_j)  for j <= i
    S_C = e^{b_C} S0 + K~^T D,         K~_j = e^{b_C - b_j} k_j
Only `tests/` may import this module.
"""
from __future__ import annotations

import numpy as np

CHUNK = 64


def chunk_prepare(q, k, v, g, beta, dtype=np.float64):
    """One chunk of one head: q, k [C, K] (q already scaled), v [C, V], g, beta [C] -> dict of W, U, Kt, Qt, P, gC."""
    q, k, v, g, beta = (np.asarray(x, dtype) for x in (q, k, v, g, beta))
    C = q.shape[0]
    b = np.cumsum(g)
    decay = np.exp(np.minimum(b[:, None] - b[None, :], 0.0))   # e^{b_i - b_j}; only j <= i is used, where the exponent is <= 0
    strict = np.tril(np.ones((C, C), bool), -1)
    incl = np.tril(np.ones((C, C), bool), 0)
    A = np.where(strict, beta[:, None] * decay * (k @ k.T), 0)
    P = np.where(incl, decay * (q @ k.T), 0)
    rhs_w = (beta * np.exp(b))[:, None] * k
    rhs_u = beta[:, None] * v
    W = np.zeros_like(rhs_w)
    U = np.zeros_like(rhs_u)
    for i in range(C):                                         # forward substitution, row by row (what the kernel does per column)
        W[i] = rhs_w[i] - A[i, :i] @ W[:i]
        U[i] = rhs_u[i] - A[i, :i] @ U[:i]
    return dict(W=W, U=U, Kt=np.exp(b[-1] - b)[:, None] * k, Qt=np.exp(b)[:, None] * q, P=P, gC=np.exp(b[-1]))


def chunk_apply(prep, S0):
    """(y [C, V], S_C [K, V], D [C, V]) of one prepared chunk from its start state S0 [K, V]."""
    D = prep["U"] - prep["W"] @ S0
    y = prep["Qt"] @ S0 + prep["P"] @ D
    return y, prep["gC"] * S0 + prep["Kt"].T @ D, D


def gated_delta_rule_chunked(q, k, v, g, beta, state, chunk=CHUNK, dtype=np.float64):
    """Same signature and result as oracle.qwen3_5.gated_delta_rule (numpy): q, k [S, Hv, K] (L2-normed, q NOT yet scaled),
    v [S, Hv, V], g, beta [S, Hv], state [Hv, K, V].  A ragged tail is padded with beta = 0, g = 0, q = k = v = 0 rows."""
    q, k, v, g, beta, state = (np.asarray(x, dtype) for x in (q, k, v, g, beta, state))
    S, Hv, K = q.shape
    V = v.shape[-1]
    q = q * dtype(1.0 / np.sqrt(K))
    n_chunks = (S + chunk - 1) // chunk
    pad = n_chunks * chunk - S
    if pad:
        z = lambda x: np.concatenate([x, np.zeros((pad,) + x.shape[1:], dtype)], 0)
        q, k, v, g, beta = z(q), z(k), z(v), z(g), z(beta)
    y = np.zeros((n_chunks * chunk, Hv, V), dtype)
    new_state = state.copy()
    for h in range(Hv):
        s = state[h].copy()
        for c in range(n_chunks):
            sl = slice(c * chunk, (c + 1) * chunk)
            prep = chunk_prepare(q[sl, h], k[sl, h], v[sl, h], g[sl, h], beta[sl, h], dtype)
            y[sl, h], s, _ = chunk_apply(prep, s)
        new_state[h] = s
    return y[:S], new_state
