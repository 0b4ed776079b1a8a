"""crane_b200 -- Python (ctypes) mirror of the reference's model interface over the C ABI.

The host language of the reference is Rust, which this image cannot build; this module is the
host-side mirror used by the tests and bench (INTEGRATION.md shows the Rust `impl ModelBackend`
a Crane maintainer would add over the same C ABI).  Names follow the reference:

    Qwen3Model            crane-core/src/models/qwen3/model.rs:34-349   (`Model`)
    Qwen3VLModel          crane-core/src/models/qwen3_5/vlm.rs:78-415   (`Qwen3_5VLModel`)

There is no CPU fallback: importing works anywhere, but constructing a model without the in-tree
CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrane_b200.so")

F32, BF16, F16 = 0, 1, 2
OK, INVALID_ARG, OOM, CUDA_ERROR, UNSUPPORTED, NOT_LOADED = 0, -1, -2, -3, -4, -5

# cb::GemmEpiMode (crane_b200/csrc/gemm.cuh)
EPI_STORE_F32, EPI_STORE_BF16, EPI_RESID_F32, EPI_SILU_MUL_BF16, EPI_GELU_ERF_BF16, EPI_GELU_TANH_BF16 = range(6)


class CraneB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"crane_b200 error {code}: {msg}")
        self.code = code


class Logits(C.Structure):
    _fields_ = [("device_ptr", C.c_void_p), ("rows", C.c_size_t), ("vocab", C.c_size_t), ("stream", C.c_void_p)]


class Sampling(C.Structure):
    """`crane_b200_sampling` (include/crane_b200.h): the per-sequence fields `sampling::sample` reads."""
    _fields_ = [("temperature", C.c_float), ("top_p", C.c_float), ("top_k", C.c_int32), ("repetition_penalty", C.c_float),
                ("frequency_penalty", C.c_float), ("presence_penalty", C.c_float), ("context", C.c_void_p), ("n_context", C.c_size_t),
                ("uniforms", C.c_void_p), ("seed", C.c_uint64)]


def make_sampling(temperature=0.0, top_p=0.0, top_k=0, repetition_penalty=1.0, frequency_penalty=0.0, presence_penalty=0.0, context=(),
                  uniforms=None, seed=0):
    """-> (Sampling, keep-alive arrays)."""
    ctx = np.ascontiguousarray(list(context), dtype=np.uint32)
    uni = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float32)
    s = Sampling(float(temperature or 0.0), float(top_p or 0.0), int(top_k or 0), float(repetition_penalty), float(frequency_penalty),
                 float(presence_penalty), ctx.ctypes.data if ctx.size else None, ctx.size, None if uni is None else uni.ctypes.data, int(seed))
    return s, (ctx, uni)


_lib = None

# name -> (restype, argtypes): must list every symbol include/crane_b200.h declares
_SIGNATURES = {
    "crane_b200_create": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "crane_b200_destroy": (None, [C.c_void_p]),
    "crane_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "crane_b200_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "crane_b200_load_tensor_ggml": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_size_t]),
    "crane_b200_load_safetensors": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "crane_b200_load_gguf": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "crane_b200_gguf_config": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "crane_b200_finalize": (C.c_int, [C.c_void_p]),
    "crane_b200_forward_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(Logits)]),
    "crane_b200_forward_step_argmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32)]),
    "crane_b200_clear_kv_cache": (C.c_int, [C.c_void_p]),
    "crane_b200_num_layers": (C.c_int, [C.c_void_p]),
    "crane_b200_warmup": (C.c_int, [C.c_void_p]),
    "crane_b200_active_kv_cache_bytes": (C.c_uint64, [C.c_void_p]),
    "crane_b200_kv_len": (C.c_size_t, [C.c_void_p]),
    "crane_b200_vocab_size": (C.c_int, [C.c_void_p]),
    "crane_b200_hidden_size": (C.c_int, [C.c_void_p]),
    "crane_b200_copy_logits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "crane_b200_forward_embeds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Logits)]),
    "crane_b200_decode_greedy": (C.c_int, [C.c_void_p, C.c_uint32, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.POINTER(C.c_size_t)]),
    "crane_b200_generate_greedy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                             C.c_void_p, C.POINTER(C.c_size_t)]),
    "crane_b200_seq_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "crane_b200_kv_export": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "crane_b200_kv_import": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "crane_b200_kv_set_len": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32]),
    "crane_b200_seq_fork": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "crane_b200_seq_free": (C.c_int, [C.c_void_p, C.c_int]),
    "crane_b200_seq_select": (C.c_int, [C.c_void_p, C.c_int]),
    "crane_b200_decode_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "crane_b200_comm_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "crane_b200_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "crane_b200_comm_world": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "crane_b200_decode_batch_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(Logits)]),
    "crane_b200_copy_gathered_logits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "crane_b200_sample": (C.c_int, [C.c_void_p, C.POINTER(Sampling), C.POINTER(C.c_uint32)]),
    "crane_b200_forward_step_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(Sampling), C.POINTER(C.c_uint32)]),
    "crane_b200_decode_batch_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Sampling), C.c_void_p]),
    "crane_b200_topk": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "crane_b200_op_qlinear": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p,
                                        C.c_float, C.c_void_p]),
    "crane_b200_op_topk": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "crane_b200_op_sample": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(Sampling), C.POINTER(C.c_uint32), C.c_void_p]),
    "crane_b200_encode_images": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "crane_b200_vl_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                        C.POINTER(Logits)]),
    "crane_b200_vl_decode_step": (C.c_int, [C.c_void_p, C.c_uint32, C.c_size_t, C.POINTER(Logits)]),
    "crane_b200_tts_text_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "crane_b200_tts_codec_embed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "crane_b200_tts_prefill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "crane_b200_tts_generate": (C.c_int, [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p]),
    "crane_b200_vl_decode_step_argmax": (C.c_int, [C.c_void_p, C.c_uint32, C.c_size_t, C.POINTER(C.c_uint32)]),
    "crane_b200_next_mrope_pos": (C.c_uint32, [C.c_void_p]),
    "crane_b200_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_size_t)]),
    "crane_b200_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "crane_b200_decode_path": (C.c_int, [C.c_void_p]),
    "crane_b200_prof_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "crane_b200_prof_report": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "crane_b200_op_gemm": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int]),
}


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and ships the 128 bytes to the other ranks)."""
    lib = load_library()
    buf = (C.c_uint8 * 128)()
    rc = lib.crane_b200_comm_unique_id(buf, 128)
    if rc != OK:
        raise CraneB200Error(rc, (lib.crane_b200_last_error(None) or b"").decode("utf-8", "replace"))
    return bytes(buf)


def shard_sequences(n_total: int, world: int, rank: int):
    """Sequence indices rank `rank` of `world` decodes: contiguous blocks, equal sizes (the all-gather needs the same n on every
    rank, so n_total must divide by world -- the caller pads its batch with idle sequences otherwise)."""
    if n_total % world:
        raise ValueError(f"{n_total} sequences do not divide over {world} ranks")
    per = n_total // world
    return list(range(rank * per, (rank + 1) * per))


def load_library():
    """dlopen the in-tree CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(crane_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """Thin owner of one `crane_b200_model*`."""

    def __init__(self, config: dict, device: int = 0, **engine_opts):
        self.lib = load_library()
        cfg = dict(config)
        if engine_opts:
            cfg["engine"] = dict(cfg.get("engine", {}), **engine_opts)
        self.config = cfg
        h = C.c_void_p()
        rc = self.lib.crane_b200_create(json.dumps(cfg).encode(), device, C.byref(h))
        if rc != OK:
            raise CraneB200Error(rc, (self.lib.crane_b200_last_error(None) or b"").decode("utf-8", "replace"))
        self.h = h
        self.vocab = self.lib.crane_b200_vocab_size(h)
        self.hidden = self.lib.crane_b200_hidden_size(h)

    def _ck(self, rc: int):
        if rc != OK:
            raise CraneB200Error(rc, (self.lib.crane_b200_last_error(self.h) or b"").decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.crane_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ----
    def load_tensor(self, name: str, arr: np.ndarray):
        """f32 arrays are rounded to bf16 for the matrices; uint16 arrays are taken as bf16 bit patterns."""
        a = np.ascontiguousarray(arr)
        if a.dtype == np.float32:
            dt = F32
        elif a.dtype == np.uint16:
            dt = BF16
        elif a.dtype == np.float16:
            dt = F16
        else:
            raise TypeError(f"{name}: unsupported dtype {a.dtype}")
        shape = (C.c_int64 * a.ndim)(*a.shape)
        self._ck(self.lib.crane_b200_load_tensor(self.h, name.encode(), dt, shape, a.ndim, _ptr(a)))

    def load_safetensors(self, path: str):
        """Register every tensor of a .safetensors file (native reader); returns (loaded, skipped)."""
        a, s = C.c_size_t(0), C.c_size_t(0)
        self._ck(self.lib.crane_b200_load_safetensors(self.h, os.fsencode(path), C.byref(a), C.byref(s)))
        return a.value, s.value

    def load_gguf(self, path: str):
        """Register every tensor of a GGUF file (native reader, quantised blocks kept as they are); returns (loaded, skipped)."""
        a, s = C.c_size_t(0), C.c_size_t(0)
        self._ck(self.lib.crane_b200_load_gguf(self.h, os.fsencode(path), C.byref(a), C.byref(s)))
        return a.value, s.value

    def load_tensor_ggml(self, name: str, ggml_type: int, shape, raw: np.ndarray):
        """Raw ggml blocks (uint8) of a [rows, cols] tensor; ggml_type 8 = Q8_0, 12 = Q4_K, 14 = Q6_K."""
        a = np.ascontiguousarray(raw, dtype=np.uint8)
        shp = (C.c_int64 * len(shape))(*shape)
        self._ck(self.lib.crane_b200_load_tensor_ggml(self.h, name.encode(), ggml_type, shp, len(shape), _ptr(a), a.size))

    def load_checkpoint(self, tensors):
        for name, arr in tensors:
            self.load_tensor(name, arr)
        self.finalize()

    def finalize(self):
        self._ck(self.lib.crane_b200_finalize(self.h))

    # ---- ModelBackend surface ----
    def forward_step(self, ids, start_pos: int) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        lg = Logits()
        self._ck(self.lib.crane_b200_forward_step(self.h, _ptr(ids), ids.size, start_pos, C.byref(lg)))
        return self.copy_logits()

    def forward_step_argmax(self, ids, start_pos: int) -> int:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        tok = C.c_uint32()
        self._ck(self.lib.crane_b200_forward_step_argmax(self.h, _ptr(ids), ids.size, start_pos, C.byref(tok)))
        return int(tok.value)

    def copy_logits(self) -> np.ndarray:
        out = np.empty(self.vocab, dtype=np.float32)
        self._ck(self.lib.crane_b200_copy_logits(self.h, _ptr(out), out.size))
        return out

    def clear_kv_cache(self):
        self._ck(self.lib.crane_b200_clear_kv_cache(self.h))

    def num_layers(self) -> int:
        return self.lib.crane_b200_num_layers(self.h)

    def kv_len(self) -> int:
        return int(self.lib.crane_b200_kv_len(self.h))

    def warmup(self):
        self._ck(self.lib.crane_b200_warmup(self.h))

    def forward_embeds(self, embeds, start_pos: int, position_ids=None) -> np.ndarray:
        e = np.ascontiguousarray(embeds, dtype=np.float32)
        p = None if position_ids is None else np.ascontiguousarray(position_ids, dtype=np.uint32)
        lg = Logits()
        self._ck(self.lib.crane_b200_forward_embeds(self.h, _ptr(e), e.shape[0], None if p is None else _ptr(p), start_pos,
                                                    C.byref(lg)))
        return self.copy_logits()

    def decode_greedy(self, first_token: int, start_pos: int, n_steps: int, eos=()):
        out = np.empty(n_steps, dtype=np.uint32)
        eos_a = np.ascontiguousarray(list(eos), dtype=np.uint32)
        n = C.c_size_t()
        self._ck(self.lib.crane_b200_decode_greedy(self.h, first_token, start_pos, n_steps,
                                                   _ptr(eos_a) if eos_a.size else None, eos_a.size, _ptr(out), C.byref(n)))
        return out[: n.value].copy()

    def generate_greedy(self, prompt, max_new_tokens: int, eos=()):
        p = np.ascontiguousarray(prompt, dtype=np.uint32)
        out = np.empty(max_new_tokens, dtype=np.uint32)
        eos_a = np.ascontiguousarray(list(eos), dtype=np.uint32)
        n = C.c_size_t()
        self._ck(self.lib.crane_b200_generate_greedy(self.h, _ptr(p), p.size, max_new_tokens,
                                                     _ptr(eos_a) if eos_a.size else None, eos_a.size, _ptr(out), C.byref(n)))
        return out[: n.value].copy()

    # ---- sequence slots + batched decode ----
    def seq_create(self) -> int:
        s = C.c_int()
        self._ck(self.lib.crane_b200_seq_create(self.h, C.byref(s)))
        return int(s.value)

    # ---- multi-GPU: sequences sharded over ranks, NCCL all-gather of each round's results ----
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * len(unique_id)).from_buffer_copy(unique_id)
        self._ck(self.lib.crane_b200_comm_init(self.h, buf, len(unique_id), rank, world))
        self._world = world

    def decode_batch_gather(self, seqs, tokens, want_logits: bool = False):
        """One decode round of this rank's sequences, then the all-gather: (tokens [world * n], logits [world * n, V] or None)."""
        seqs = np.ascontiguousarray(seqs, dtype=np.int32)
        tokens = np.ascontiguousarray(tokens, dtype=np.uint32)
        world = getattr(self, "_world", 1)
        out = np.empty(world * seqs.size, dtype=np.uint32)
        lg = Logits()
        self._ck(self.lib.crane_b200_decode_batch_gather(self.h, _ptr(seqs), _ptr(tokens), seqs.size, _ptr(out), C.byref(lg) if want_logits else None))
        if not want_logits:
            return out, None
        rows, V = int(lg.rows), int(lg.vocab)
        host = np.empty((rows, V), dtype=np.float32)
        self._ck(self.lib.crane_b200_copy_gathered_logits(self.h, _ptr(host), host.size))
        return out, host

    # ---- KV swap: ModelBackend::get_kv_caches / set_kv_caches ----
    def get_kv_caches(self):
        """[(K, V)] per layer for the current sequence: attention layers [n_kv, T, D] f32; GDN layers (conv window, recurrent state)."""
        from . import synth
        tc = synth.text_config(self.config)
        out = []
        for layer in range(self.num_layers()):
            n = C.c_size_t()
            self._ck(self.lib.crane_b200_kv_export(self.h, layer, None, None, 0, C.byref(n)))
            if "linear_num_value_heads" in tc and not synth.is_full_attention_layer(tc, layer):
                conv_dim = 2 * tc["linear_num_key_heads"] * tc["linear_key_head_dim"] + tc["linear_num_value_heads"] * tc["linear_value_head_dim"]
                k = np.empty((conv_dim, tc.get("linear_conv_kernel_dim", 4)), np.float32)
                v = np.empty((tc["linear_num_value_heads"], tc["linear_key_head_dim"], tc["linear_value_head_dim"]), np.float32)
                cap = max(k.size, v.size)
            else:
                nkv, D = tc["num_key_value_heads"], synth.head_dim(tc)
                k = np.empty((nkv, n.value, D), np.float32)
                v = np.empty_like(k)
                cap = k.size
            if cap:
                self._ck(self.lib.crane_b200_kv_export(self.h, layer, _ptr(k), _ptr(v), cap, C.byref(n)))
            out.append((k, v))
        return out

    def set_kv_caches(self, caches, n_tokens: int, next_rotary_pos: int = None):
        for layer, (k, v) in enumerate(caches):
            k = np.ascontiguousarray(k, dtype=np.float32)
            v = np.ascontiguousarray(v, dtype=np.float32)
            self._ck(self.lib.crane_b200_kv_import(self.h, layer, _ptr(k), _ptr(v), n_tokens))
        self._ck(self.lib.crane_b200_kv_set_len(self.h, n_tokens, n_tokens if next_rotary_pos is None else next_rotary_pos))

    def seq_fork(self, src: int) -> int:
        """A new sequence that starts as a copy of `src` (KV pages, GDN state, length, rotary position)."""
        s = C.c_int()
        self._ck(self.lib.crane_b200_seq_fork(self.h, src, C.byref(s)))
        return int(s.value)

    def seq_free(self, seq: int):
        self._ck(self.lib.crane_b200_seq_free(self.h, seq))

    def seq_select(self, seq: int):
        self._ck(self.lib.crane_b200_seq_select(self.h, seq))

    def decode_batch(self, seqs, tokens, n_steps: int = 1, want_logits: bool = False):
        """`step_batch_decode` x n_steps for the listed sequences -> (tokens [n, n_steps], logits [n, V] of the last round | None)."""
        sq = np.ascontiguousarray(seqs, dtype=np.int32)
        tk = np.ascontiguousarray(tokens, dtype=np.uint32)
        out = np.empty((sq.size, n_steps), dtype=np.uint32)
        lg = np.empty((sq.size, self.vocab), dtype=np.float32) if want_logits else None
        self._ck(self.lib.crane_b200_decode_batch(self.h, _ptr(sq), _ptr(tk), sq.size, n_steps, _ptr(out), None if lg is None else _ptr(lg)))
        return out, lg

    # ---- device-side sampling (crane-serve/src/engine/sampling.rs) ----
    def sample(self, **kw) -> int:
        """`sampling::sample` on the logits of the last forward call (they never leave the device)."""
        sp, keep = make_sampling(**kw)
        tok = C.c_uint32()
        self._ck(self.lib.crane_b200_sample(self.h, C.byref(sp), C.byref(tok)))
        return int(tok.value)

    def forward_step_sample(self, ids, start_pos: int, **kw) -> int:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        sp, keep = make_sampling(**kw)
        tok = C.c_uint32()
        self._ck(self.lib.crane_b200_forward_step_sample(self.h, _ptr(ids), ids.size, start_pos, C.byref(sp), C.byref(tok)))
        return int(tok.value)

    def decode_batch_sample(self, seqs, tokens, params):
        """One decode round with per-sequence sampling; params = list of make_sampling keyword dicts."""
        sq = np.ascontiguousarray(seqs, dtype=np.int32)
        tk = np.ascontiguousarray(tokens, dtype=np.uint32)
        made = [make_sampling(**p) for p in params]
        arr = (Sampling * len(made))(*[m[0] for m in made])
        out = np.empty(sq.size, dtype=np.uint32)
        self._ck(self.lib.crane_b200_decode_batch_sample(self.h, _ptr(sq), _ptr(tk), sq.size, arr, _ptr(out)))
        return out

    def topk(self, k: int):
        idx = np.empty(k, dtype=np.uint32)
        val = np.empty(k, dtype=np.float32)
        self._ck(self.lib.crane_b200_topk(self.h, k, _ptr(idx), _ptr(val)))
        return idx, val

    # ---- vision-language surface ----
    def encode_images(self, pixel_values, grid_thw, want_deepstack: int = 0):
        pv = np.ascontiguousarray(pixel_values, dtype=np.float32)
        g = np.ascontiguousarray(grid_thw, dtype=np.uint32).reshape(-1, 3)
        merge = self.config["vision_config"]["spatial_merge_size"]
        n_tok = int(sum(int(t) * int(h) * int(w) for t, h, w in g) // (merge * merge))
        out = np.empty((n_tok, self.hidden), dtype=np.float32)
        ds = np.empty((want_deepstack, n_tok, self.hidden), dtype=np.float32) if want_deepstack else None
        self._ck(self.lib.crane_b200_encode_images(self.h, _ptr(pv), _ptr(g), g.shape[0], _ptr(out), None if ds is None else _ptr(ds)))
        return out, ds

    def vl_forward(self, ids, pixel_values=None, grid_thw=None, start_pos: int = 0) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        lg = Logits()
        if pixel_values is None:
            self._ck(self.lib.crane_b200_vl_forward(self.h, _ptr(ids), ids.size, None, None, 0, start_pos, C.byref(lg)))
        else:
            pv = np.ascontiguousarray(pixel_values, dtype=np.float32)
            g = np.ascontiguousarray(grid_thw, dtype=np.uint32).reshape(-1, 3)
            self._ck(self.lib.crane_b200_vl_forward(self.h, _ptr(ids), ids.size, _ptr(pv), _ptr(g), g.shape[0], start_pos, C.byref(lg)))
        return self.copy_logits()

    def vl_decode_step(self, token: int, start_pos: int) -> np.ndarray:
        lg = Logits()
        self._ck(self.lib.crane_b200_vl_decode_step(self.h, token, start_pos, C.byref(lg)))
        return self.copy_logits()

    def vl_decode_step_argmax(self, token: int, start_pos: int) -> int:
        t = C.c_uint32()
        self._ck(self.lib.crane_b200_vl_decode_step_argmax(self.h, token, start_pos, C.byref(t)))
        return int(t.value)

    def next_mrope_pos(self) -> int:
        return int(self.lib.crane_b200_next_mrope_pos(self.h))

    def last_timing(self):
        p, d, n = C.c_float(), C.c_float(), C.c_size_t()
        self._ck(self.lib.crane_b200_last_timing(self.h, C.byref(p), C.byref(d), C.byref(n)))
        return {"prefill_ms": p.value, "decode_ms": d.value, "decode_steps": int(n.value)}

    def active_kv_cache_bytes(self) -> int:
        return int(self.lib.crane_b200_active_kv_cache_bytes(self.h))

    def kernel_launches(self) -> int:
        return int(self.lib.crane_b200_kernel_launches(self.h))

    def prof_enable(self, on: bool = True):
        self._ck(self.lib.crane_b200_prof_enable(self.h, 1 if on else 0))

    def prof_report(self) -> dict:
        """Per-pass averages since prof_enable: {"decode" | "prefill": {passes, tokens, enqueue_ms, wall_ms, device_ms, spans}}."""
        import json
        n = C.c_size_t()
        self._ck(self.lib.crane_b200_prof_report(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self._ck(self.lib.crane_b200_prof_report(self.h, buf, n.value, C.byref(n)))
        return json.loads(buf.value.decode())

    def decode_path(self) -> str:
        return "persistent" if self.lib.crane_b200_decode_path(self.h) else "chain"


class Qwen3Model(Engine):
    """`crane_core::models::qwen3::Model` (qwen3/model.rs:34-349) without the tokenizer."""

    @classmethod
    def from_gguf(cls, path: str, device: int = 0, **engine_opts):
        """`Model::new_with_format(.., Gguf, ..)` -> `Qwen3Model::from_gguf` (qwen3/modeling.rs:821-935): config from the file's
        metadata, tensors (quantised blocks included) from the file itself."""
        m = cls(gguf_config(path), device=device, **engine_opts)
        m.load_gguf(path)
        m.finalize()
        return m

    def generate(self, input_ids, max_new_tokens: int = 128, eos_token_id=(), temperature=None):
        """`ModelForCausalLM::generate` (generation/based.rs:5-34); greedy only (temperature None => ArgMax,
        qwen3/model.rs:281-284).  Sampling stays on the caller's side of the boundary."""
        if temperature not in (None, 0, 0.0):
            raise CraneB200Error(UNSUPPORTED, "sampled decoding is host-side in the reference (LogitsProcessor); only greedy here")
        eos = (eos_token_id,) if isinstance(eos_token_id, int) else tuple(eos_token_id)
        return self.generate_greedy(input_ids, max_new_tokens, eos)


class Qwen3_5Model(Qwen3Model):
    """`crane_core::models::qwen3_5::Model` (qwen3_5/model.rs:628-943): hybrid Gated-Delta-Net + gated attention
    decoder.  Same surface as Qwen3Model; `clear_kv_cache` also zeroes the GDN conv / recurrent state."""


class Qwen3VLModel(Engine):
    """`Qwen3_5VLModel` (qwen3_5/vlm.rs:78-415) over the Qwen3-VL geometry: `forward` = ViT + splice + prefill,
    `decode_step` = one token with the scalar MRoPE counter."""

    def forward(self, input_ids, pixel_values=None, image_grid_thw=None, start_pos: int = 0):
        return self.vl_forward(input_ids, pixel_values, image_grid_thw, start_pos)

    def decode_step(self, token: int, start_pos: int):
        return self.vl_decode_step(token, start_pos)

    def decode_step_argmax(self, token: int, start_pos: int) -> int:
        return self.vl_decode_step_argmax(token, start_pos)

    def generate(self, input_ids, pixel_values, image_grid_thw, max_new_tokens: int, eos_token_id=()):
        """vlm.rs:355-414: prefill (argmax on device), then greedy decode on the device."""
        eos = (eos_token_id,) if isinstance(eos_token_id, int) else tuple(eos_token_id)
        self.clear_kv_cache()
        logits = self.forward(input_ids, pixel_values, image_grid_thw, 0)
        first = int(np.argmax(logits))
        out = [first]
        if first in eos or max_new_tokens == 1:
            return np.array(out, np.uint32)
        rest = self.decode_greedy(first, len(input_ids), max_new_tokens - 1, eos)
        return np.concatenate([np.array(out, np.uint32), rest])


class Qwen3TTSModel(Engine):
    """`Qwen3TTSModel` codec-LM (crane-core/src/models/qwen3_tts/modeling.rs:1347-1760): talker + code predictor.
    The host glue (`build_prefill_embeds`, :597-726) lives here exactly as it lives in Rust in the reference; the
    arithmetic (text projection, embedding gathers, every transformer pass, the frame loop) runs behind the C ABI."""

    def __init__(self, config: dict, device: int = 0, **engine_opts):
        super().__init__(config, device, **engine_opts)
        self.tk = config["talker_config"]
        self.groups = self.tk["num_code_groups"]
        self.cp_vocab = self.tk["code_predictor_config"]["vocab_size"]

    def text_project(self, ids) -> np.ndarray:
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.empty((a.size, self.hidden), np.float32)
        self._ck(self.lib.crane_b200_tts_text_project(self.h, _ptr(a), a.size, _ptr(out)))
        return out

    def codec_embed(self, ids, group: int = -1) -> np.ndarray:
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.empty((a.size, self.hidden), np.float32)
        self._ck(self.lib.crane_b200_tts_codec_embed(self.h, group, _ptr(a), a.size, _ptr(out)))
        return out

    def build_prefill_embeds(self, text_ids, language_id=None, speaker_id=None):
        """TalkerModel::build_prefill_embeds (modeling.rs:597-726) -> (prefill [P, H], trailing_text [n, H], tts_pad [H])."""
        tk, cfg = self.tk, self.config
        vt = tk["text_vocab_size"]
        role = self.text_project([151644 % vt, 77091 % vt, 198])
        tts = self.text_project([cfg["tts_pad_token_id"], cfg["tts_bos_token_id"], cfg["tts_eos_token_id"]])
        pad, bos, eos = tts[0], tts[1], tts[2]
        if language_id is not None:
            codec = [tk["codec_think_id"], tk["codec_think_bos_id"], language_id, tk["codec_think_eos_id"]]
        else:
            codec = [tk["codec_nothink_id"], tk["codec_think_bos_id"], tk["codec_think_eos_id"]]
        if speaker_id is not None:
            codec.append(speaker_id)
        codec += [tk["codec_pad_id"], tk["codec_bos_id"]]
        ce = self.codec_embed(codec)
        n_over = len(codec) - 1
        overlay = np.concatenate([np.repeat(pad[None], n_over - 1, 0), bos[None]], 0)
        codec_hidden = overlay + ce[:n_over]
        first = (self.text_project(text_ids[:1])[0] if len(text_ids) else pad) + ce[-1]
        prefill = np.concatenate([role, codec_hidden, first[None]], 0).astype(np.float32)
        trailing = np.concatenate([self.text_project(text_ids[1:]), eos[None]], 0) if len(text_ids) > 1 else eos[None]
        return prefill, trailing.astype(np.float32), pad

    def generate_codes(self, text_ids, max_new_tokens: int, repetition_penalty: float = 1.0, forced_frames=None, want_logits=False):
        """`generate_speech_codes` (modeling.rs:1429-1596), greedy or teacher-forced -> frames [n, groups] (+ logits)."""
        self.clear_kv_cache()
        prefill, trailing, pad = self.build_prefill_embeds(list(text_ids))
        prefill = np.ascontiguousarray(prefill); trailing = np.ascontiguousarray(trailing); pad = np.ascontiguousarray(pad)
        self._ck(self.lib.crane_b200_tts_prefill(self.h, _ptr(prefill), prefill.shape[0], _ptr(trailing), trailing.shape[0], _ptr(pad)))
        frames = np.zeros((max_new_tokens, self.groups), np.uint32)
        forced = None if forced_frames is None else np.ascontiguousarray(
            np.concatenate([np.asarray(forced_frames, np.uint32).reshape(-1, self.groups),
                            np.zeros((max_new_tokens - len(forced_frames), self.groups), np.uint32)], 0))
        fl = np.zeros((max_new_tokens, self.vocab), np.float32) if want_logits else None
        gl = np.zeros((max_new_tokens, self.groups - 1, self.cp_vocab), np.float32) if want_logits else None
        n = C.c_size_t()
        self._ck(self.lib.crane_b200_tts_generate(self.h, max_new_tokens, repetition_penalty, None if forced is None else _ptr(forced),
                                                  _ptr(frames), C.byref(n), None if fl is None else _ptr(fl), None if gl is None else _ptr(gl)))
        k = int(n.value)
        return (frames[:k], fl[:k], gl[:k]) if want_logits else frames[:k]


def gguf_config(path: str) -> dict:
    """config.json of a GGUF checkpoint, derived by the library from the file's metadata (pure host code, no GPU needed)."""
    lib = load_library()
    buf = C.create_string_buffer(16384)
    need = C.c_size_t(0)
    rc = lib.crane_b200_gguf_config(os.fsencode(path), buf, len(buf), C.byref(need))
    if rc != OK and need.value > len(buf):            # a model with very many layers: the text's size comes back in `needed`
        buf = C.create_string_buffer(need.value)
        rc = lib.crane_b200_gguf_config(os.fsencode(path), buf, len(buf), C.byref(need))
    if rc != OK:
        raise CraneB200Error(rc, (lib.crane_b200_last_error(None) or b"").decode("utf-8", "replace"))
    return json.loads(buf.value.decode())


def op_gemm(a_bits: np.ndarray, w_bits: np.ndarray, mode: int, bias=None, out_init=None, use_simt=False, device=0, a_lo_bits=None):
    """Kernel-level test hook: epilogue((A + A_lo)[M,K] x W[N,K]^T) with bf16 bit-pattern inputs."""
    lib = load_library()
    M, K = a_bits.shape
    N = w_bits.shape[0]
    half = mode in (EPI_STORE_BF16, EPI_SILU_MUL_BF16, EPI_GELU_ERF_BF16, EPI_GELU_TANH_BF16)
    cols = N // 2 if mode == EPI_SILU_MUL_BF16 else N
    out = np.zeros((M, cols), dtype=np.uint16 if half else np.float32) if out_init is None else np.ascontiguousarray(out_init).copy()
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    alo = None if a_lo_bits is None else np.ascontiguousarray(a_lo_bits)
    rc = lib.crane_b200_op_gemm(device, _ptr(np.ascontiguousarray(a_bits)), None if alo is None else _ptr(alo),
                                _ptr(np.ascontiguousarray(w_bits)), M, N, K, mode,
                                None if b is None else _ptr(b), _ptr(out), 1 if use_simt else 0)
    if rc != OK:
        raise CraneB200Error(rc, (lib.crane_b200_last_error(None) or b"").decode("utf-8", "replace"))
    return out


def op_topk(logits: np.ndarray, k: int, device=0) -> np.ndarray:
    """`crane_core::ops::topk_indices` on caller logits (kernel-level test hook)."""
    lib = load_library()
    x = np.ascontiguousarray(logits, dtype=np.float32)
    out = np.empty(k, dtype=np.uint32)
    rc = lib.crane_b200_op_topk(device, _ptr(x), x.size, k, _ptr(out))
    if rc != OK:
        raise CraneB200Error(rc, (lib.crane_b200_last_error(None) or b"").decode("utf-8", "replace"))
    return out


def op_sample(logits: np.ndarray, device=0, **kw):
    """`sampling::sample` on caller logits -> (token, logits after penalties)."""
    lib = load_library()
    x = np.ascontiguousarray(logits, dtype=np.float32)
    sp, keep = make_sampling(**kw)
    tok = C.c_uint32()
    after = np.empty_like(x)
    rc = lib.crane_b200_op_sample(device, _ptr(x), x.size, C.byref(sp), C.byref(tok), _ptr(after))
    if rc != OK:
        raise CraneB200Error(rc, (lib.crane_b200_last_error(None) or b"").decode("utf-8", "replace"))
    return int(tok.value), after


def op_qlinear(x: np.ndarray, raw: np.ndarray, ggml_type: int, n: int, norm_w=None, eps: float = 1e-6, device=0) -> np.ndarray:
    """Quantised linear of x [m, k] through the decode kernels (kernel-level test hook) -> [m, n] f32."""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float32)
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    m, k = x.shape
    y = np.empty((m, n), np.float32)
    nw = None if norm_w is None else np.ascontiguousarray(norm_w, dtype=np.float32)
    rc = lib.crane_b200_op_qlinear(device, _ptr(x), m, k, _ptr(raw), raw.size, ggml_type, n, None if nw is None else _ptr(nw), eps, _ptr(y))
    if rc != OK:
        raise CraneB200Error(rc, (lib.crane_b200_last_error(None) or b"").decode("utf-8", "replace"))
    return y
