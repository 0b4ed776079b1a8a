#!/usr/bin/env python
"""Where does a quantised model's output leave the integer-dot oracle?  Three forward passes of the same tiny GGUF-quantised
model: the CPU oracle, the oracle with every quantised linear replaced by the GPU kernels (op_qlinear), and the engine."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import crane_b200  # noqa: E402
from crane_b200 import synth  # noqa: E402
from oracle import ggml_quant as gq  # noqa: E402
from oracle.qwen3 import Qwen3Oracle  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


import ctypes as C


def peek(m, which, shape):
    out = np.empty(shape, np.float32)
    fn = m.lib.crane_b200_debug_peek
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    assert fn(m.h, which, out.ctypes.data_as(C.c_void_p), out.size) == 0
    return out


class Capture(Qwen3Oracle):
    """Keeps the inputs / outputs of every linear of layer 0."""
    def _linear(self, full_name, x, w=None):
        y = super()._linear(full_name, x, w)
        self.cap = getattr(self, "cap", {})
        self.cap.setdefault(full_name, (x.numpy().copy(), y.numpy().copy()))
        return y

    def _layer(self, i, x, cos, sin, kv_offset):
        out = super()._layer(i, x, cos, sin, kv_offset)
        self.cap["x_after_layer_%d" % i] = out.numpy().copy()
        return out


class GpuLinearOracle(Qwen3Oracle):
    def _linear(self, full_name, x, w=None):
        if full_name in self.q:
            raw, qt = self.q[full_name]
            x2 = x.reshape(-1, x.shape[-1]).numpy()
            y = crane_b200.op_qlinear(x2, raw, gq.GGML_TYPE_ID[qt], raw.shape[0])
            return torch.from_numpy(y).reshape(*x.shape[:-1], -1)
        return super()._linear(full_name, x, w)


cfg = synth.TINY_QWEN3
recipe = {k: "Q8_0" for k in ("q_proj.weight", "k_proj.weight", "v_proj.weight", "o_proj.weight", "gate_proj.weight", "up_proj.weight",
                               "down_proj.weight", "embed_tokens.weight")}
for nl in (1,):
    c = dict(cfg, num_hidden_layers=nl)
    w = dict(synth.synth_checkpoint(c))
    m = crane_b200.Qwen3Model(c, device=0, max_seq_len=256)
    wq, qd = {}, {}
    for name, arr in w.items():
        qt = next((t for suf, t in recipe.items() if name.endswith(suf)), None)
        if qt is None or arr.ndim != 2:
            m.load_tensor(name, arr)
            wq[name] = arr
        else:
            raw = gq.quantize(arr, qt)
            m.load_tensor_ggml(name, gq.GGML_TYPE_ID[qt], arr.shape, raw)
            wq[name] = gq.dequantize(raw, qt, arr.shape[1])
            qd[name] = (raw, qt)
    m.finalize()
    for S in (1, 2, 3, 4, 5, 8):
        ids = synth.synth_token_ids(S, c["vocab_size"], "qdbg")
        a = Qwen3Oracle(c, wq, quantised=qd).forward(ids, 0).numpy()
        b = GpuLinearOracle(c, wq, quantised=qd).forward(ids, 0).numpy()
        m.clear_kv_cache()
        g = m.forward_step(ids, 0)
        print(f"layers {nl} S {S}: oracle-with-GPU-linears vs oracle {rel(b, a):.2e} | engine vs oracle {rel(g, a):.2e} | engine vs GPU-linears {rel(g, b):.2e}")
        if S == 1:
            continue
        cap = Capture(c, wq, quantised=qd)
        cap.forward(ids, 0)
        H, I = c["hidden_size"], c["intermediate_size"]
        qn = [cap.cap[f"model.layers.0.self_attn.{t}_proj.weight"][1] for t in "qkv"]
        qkv_ref = np.concatenate(qn, -1)
        qkv = peek(m, 1, qkv_ref.shape)
        act_ref = cap.cap["model.layers.0.mlp.down_proj.weight"][0]
        act = peek(m, 2, (S, I))
        x_ref = cap.cap["x_after_layer_0"]
        xg = peek(m, 0, (S, H))
        per_row = lambda u, v: " ".join(f"{rel(u[r], v[r]):.1e}" for r in range(S))
        print(f"    qkv rows: {per_row(qkv, qkv_ref)}\n    act rows: {per_row(act, act_ref)}\n    x   rows: {per_row(xg, x_ref)}")
    m.close()
