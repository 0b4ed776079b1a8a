"""CPU oracle for the Qwen3-TTS codec-LM frame loop (talker + code predictor) -- TEST INFRASTRUCTURE ONLY.

Restates (all crane-core/src/models/qwen3_tts/modeling.rs):
  ResizeMlp text projection ......................... :244-268
  CodePredictor::{new, predict} ..................... :277-479   (2-token prefill, then one token per remaining code group)
  TalkerModel::{build_prefill_embeds, forward_embeds} :597-747
  Qwen3TTSModel::generate_speech_codes frame loop ... :1429-1596 (suppress mask, EOS suppressed for 2 steps, repetition
                                                       penalty over first codes, trailing-text / tts_pad contribution)
over the same transformer block as the dense Qwen3 decoder (modules/transformer.rs:124-146, modules/attention.rs:207-346:
pre-norm, per-head QK-norm before half-split RoPE, SwiGLU) -- arithmetic shared with oracle/qwen3.py, which is pinned to HF.

Parity status: the decoder stack is pinned (via Qwen3Oracle vs HF); the TTS glue is restated from the reference only, and the
reference's sampler (`LogitsProcessor::from_sampling(42, TopKThenTopP{k: 50})`, candle-transformers, un-vendored) is NOT
pinned: the oracle takes a `pick(logits, kind, index)` callback (default: first maximum) or forced codes, so parity is
checked on logits given forced tokens and on greedy codes.
"""
from __future__ import annotations

import numpy as np
import torch

from .qwen3 import Qwen3Oracle, argmax_first, apply_repeat_penalty, silu


def _stack(cfg_part: dict, weights: dict, prefix: str, max_pos: int) -> Qwen3Oracle:
    """A Qwen3Oracle over `prefix`layers.* / `prefix`norm.weight (embedding / head unused)."""
    w = {k: v for k, v in weights.items() if k.startswith(prefix)}
    H = cfg_part["hidden_size"]
    w[prefix + "embed_tokens.weight"] = np.zeros((2, H), np.float32)
    cfg = dict(cfg_part, vocab_size=2, tie_word_embeddings=True)
    return Qwen3Oracle(cfg, w, prefix=prefix, max_pos=max_pos)


class Qwen3TTSOracle:
    def __init__(self, cfg: dict, weights: dict, max_pos: int = 4096):
        self.cfg = cfg
        self.tk = cfg["talker_config"]
        self.cp = self.tk["code_predictor_config"]
        self.w = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).float() for k, v in weights.items()}
        self.talker = _stack(self.tk, weights, "talker.model.", max_pos)
        self.pred = _stack(self.cp, weights, "talker.code_predictor.model.", max_pos)
        self.n_groups = self.cp["num_code_groups"]
        self.eos = int(self.tk["codec_eos_token_id"])
        self.Vc = int(self.tk["vocab_size"])

    # ---- embeddings ----------------------------------------------------------------------------------------------
    def text_project(self, ids) -> torch.Tensor:
        """text_embedding -> ResizeMlp (fc1 + bias, SiLU, fc2 + bias)   (:244-268, :612-614)"""
        e = self.w["talker.model.text_embedding.weight"][torch.as_tensor(np.asarray(ids, np.int64))]
        h = silu(e @ self.w["talker.text_projection.linear_fc1.weight"].T + self.w["talker.text_projection.linear_fc1.bias"])
        return h @ self.w["talker.text_projection.linear_fc2.weight"].T + self.w["talker.text_projection.linear_fc2.bias"]

    def codec_embed(self, ids) -> torch.Tensor:
        return self.w["talker.model.codec_embedding.weight"][torch.as_tensor(np.asarray(ids, np.int64))]

    def group_embed(self, g: int, ids) -> torch.Tensor:
        return self.w[f"talker.code_predictor.model.codec_embedding.{g}.weight"][torch.as_tensor(np.asarray(ids, np.int64))]

    def build_prefill_embeds(self, text_ids, language_id=None, speaker_id=None):
        """TalkerModel::build_prefill_embeds (:597-726) -> (prefill [P, H], trailing_text [n, H], tts_pad [H])."""
        tk = self.tk
        role = self.text_project([151644 % tk["text_vocab_size"], 77091 % tk["text_vocab_size"], 198])
        tts = self.text_project([self.cfg["tts_pad_token_id"], self.cfg["tts_bos_token_id"], self.cfg["tts_eos_token_id"]])
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
        overlay = torch.cat([pad[None].expand(n_over - 1, -1), bos[None]], 0)
        codec_hidden = overlay + ce[:n_over]
        first = (self.text_project(text_ids[:1])[0] if len(text_ids) else pad) + ce[-1]
        prefill = torch.cat([role, codec_hidden, first[None]], 0)
        trailing = torch.cat([self.text_project(text_ids[1:]), eos[None]], 0) if len(text_ids) > 1 else eos[None]
        return prefill, trailing, pad

    # ---- code predictor (:373-479) ---------------------------------------------------------------------------------
    def predict_codes(self, talker_hidden: torch.Tensor, first_code: int, pick=None, forced=None):
        """-> (codes [n_groups-1], logits [n_groups-1, Vcp]).  `forced`: teacher-forced codes fed back instead of the picks."""
        pick = pick or (lambda lg, kind, i: argmax_first(lg))
        self.pred.clear_kv_cache()
        proj_w = self.w.get("talker.code_predictor.small_to_mtp_projection.weight")

        def proj(x):
            return x if proj_w is None else x @ proj_w.T + self.w["talker.code_predictor.small_to_mtp_projection.bias"]
        x = torch.stack([talker_hidden, self.codec_embed([first_code])[0]], 0)
        self.pred.forward_embeds(proj(x), 0)
        codes, logits = [], []
        for g in range(self.n_groups - 1):
            if g > 0:
                e = self.group_embed(g - 1, [codes[-1]])
                self.pred.forward_embeds(proj(e), 1 + g)
            hid = self.pred.last_hidden_states[-1]
            lg = hid @ self.w[f"talker.code_predictor.lm_head.{g}.weight"].T
            logits.append(lg)
            c = int(pick(lg, "group", g))
            codes.append(int(forced[g]) if forced is not None else c)
        self.pred.clear_kv_cache()
        return codes, torch.stack(logits)

    # ---- frame loop (:1429-1596) ---------------------------------------------------------------------------------------
    def first_code_logits(self, past_hidden, step, history, repetition_penalty=1.0):
        lg = (past_hidden @ self.w["talker.codec_head.weight"].T).numpy()
        if repetition_penalty != 1.0 and history:
            lg = apply_repeat_penalty(lg, repetition_penalty, [c[0] for c in history])
        sup = np.zeros(self.Vc, np.float32)
        start = max(0, self.Vc - 1024)
        sup[start:] = -np.inf
        sup[self.eos] = 0.0
        lg = lg + sup
        if step < 2:
            lg[self.eos] = -np.inf
        return torch.from_numpy(lg.astype(np.float32))

    def generate_codes(self, text_ids, max_new_tokens, repetition_penalty=1.0, pick=None, forced_frames=None):
        """-> (frames [n, n_groups], per-frame dict of logits).  Greedy by default; `forced_frames` teacher-forces every code."""
        pick = pick or (lambda lg, kind, i: argmax_first(lg))
        self.talker.clear_kv_cache()
        prefill, trailing, pad = self.build_prefill_embeds(text_ids)
        P = prefill.shape[0]
        self.talker.forward_embeds(prefill, 0)
        past = self.talker.last_hidden_states[-1]
        frames, trace = [], []
        for step in range(max_new_tokens):
            lg0 = self.first_code_logits(past, step, frames, repetition_penalty)
            first = int(forced_frames[step][0]) if forced_frames is not None else int(pick(lg0, "first", step))
            if first == self.eos:
                break
            codes, lgs = self.predict_codes(past, first, pick, None if forced_frames is None else forced_frames[step][1:])
            frame = [first] + codes
            frames.append(frame)
            trace.append({"first_logits": lg0, "group_logits": lgs, "hidden": past})
            emb = self.codec_embed([first])[0]
            for i, c in enumerate(codes):
                emb = emb + self.group_embed(i, [c])[0]
            emb = emb + (trailing[step] if step < trailing.shape[0] else pad)
            self.talker.forward_embeds(emb[None], P + step)
            past = self.talker.last_hidden_states[-1]
        return frames, trace
