"""CPU: the oracle against (a) the committed HF-generated fixtures and (b) every literal known answer /
property the reference's own tests hold for this path."""
import math

import numpy as np
import torch

from conftest import golden, rel_err
from crane_b200 import synth
from oracle import qwen3 as oq
from oracle import qwen3_vl as ov


def _weights(cfg):
    return dict(synth.synth_checkpoint(cfg))


def test_tiny_qwen3_matches_hf_fixture():
    for name, cfg in (("tiny_qwen3", synth.TINY_QWEN3), ("tiny_qwen3_untied", synth.TINY_QWEN3_UNTIED)):
        g = golden(name)
        orc = oq.Qwen3Oracle(cfg, _weights(cfg))
        toks = [int(t) for t in g["prompt"]]
        for step in range(g["logits"].shape[0]):
            ctx = toks if step == 0 else toks[-1:]
            lo = orc.forward(ctx, len(toks) - len(ctx)).numpy()
            assert rel_err(lo, g["logits"][step]) < 2e-5
            assert oq.argmax_first(torch.from_numpy(lo)) == int(g["tokens"][step])
            toks.append(int(g["tokens"][step]))


def test_tiny_qwen3_vl_matches_hf_fixture():
    cfg = synth.TINY_QWEN3_VL
    g = golden("tiny_qwen3_vl")
    w = _weights(cfg)
    pv, grid = synth.patchify(g["image"])
    assert tuple(grid) == tuple(int(x) for x in g["grid"])
    hf = ov.Qwen3VLOracle(cfg, w, vit_act="tanh", merger_act="erf")      # HF's activation choice
    img, _ = hf.vision.forward(pv, [grid])
    assert rel_err(img.numpy(), g["hf_image_embeds"]) < 2e-5
    lo = hf.prefill(g["prompt"], pv, [grid]).numpy()
    assert rel_err(lo, g["hf_logits"][0]) < 2e-5
    S = len(g["prompt"])
    tok = int(g["hf_tokens"][0])
    for step in range(1, g["hf_logits"].shape[0]):
        lo = hf.decode_step(tok, S + step - 1).numpy()
        assert rel_err(lo, g["hf_logits"][step]) < 2e-5
        tok = int(g["hf_tokens"][step])
    ref = ov.Qwen3VLOracle(cfg, w)                                          # the reference's activation choice
    img, deep = ref.vision.forward(pv, [grid])
    assert rel_err(img.numpy(), g["ref_image_embeds"]) < 1e-6
    assert rel_err(torch.stack(deep).numpy(), g["ref_deepstack"]) < 1e-6
    assert rel_err(ref.prefill(g["prompt"], pv, [grid]).numpy(), g["ref_logits"][0]) < 1e-6


# ---- literal known answers from the reference's tests ------------------------------------------------

def test_rope_inv_freq_literals():
    # crane-core/src/models/modules/rotary.rs:166-189: dim=8, theta=1e4 => inv_freq [1, .1, .01, .001]
    cos, sin = oq.rope_tables(8, 2, 10000.0)
    for i, f in enumerate([1.0, 0.1, 0.01, 0.001]):
        assert abs(float(cos[1, i]) - math.cos(f)) < 1e-5 and abs(float(sin[1, i]) - math.sin(f)) < 1e-5
    # rotary.rs:214-235: dim=4, theta=100 => inv_freq [1, .1] at several positions
    cos, sin = oq.rope_tables(4, 16, 100.0)
    for pos in (0, 1, 5, 10):
        for i, f in enumerate([pos * 1.0, pos * 0.1]):
            assert abs(float(cos[pos, i]) - math.cos(f)) < 1e-5 and abs(float(sin[pos, i]) - math.sin(f)) < 1e-5


def test_topk_order_literals():
    # crane-core/tests/rocm_kernels.rs:169-172 and the tie rule :141-160 (value desc, index asc)
    assert list(oq.topk_order(np.array([0.5, -3, 7.25, 1, 7.5], np.float32), 5)) == [4, 2, 3, 0, 1]
    v = np.zeros(64, np.float32)
    v[3::4] = 1.0
    assert list(oq.topk_order(v, 5)) == [3, 7, 11, 15, 19]


def test_causal_mask_rows():
    # crane-core/src/models/qwen3_5/prefill.rs:283-296 and qwen3/modeling.rs:1000-1014
    assert oq.build_causal_mask_rows(3, 0) == [[1, 0, 0], [1, 1, 0], [1, 1, 1]]
    assert oq.build_causal_mask_rows(2, 3) == [[1, 1, 1, 1, 0], [1, 1, 1, 1, 1]]


def test_repeat_penalty_rule():
    # crane-core/src/models/utils.rs:25-44
    out = oq.apply_repeat_penalty(np.array([2.0, -2.0, 1.0], np.float32), 2.0, [0, 1, 1, 0])
    assert list(out) == [1.0, -4.0, 1.0]


def test_patch_order_literal():
    # crane-core/src/models/qwen3_5/processor.rs:306-314: a 4x2-patch image emits patches in order [0,1,4,5,2,3,6,7]
    img = np.zeros((32, 64, 3), np.uint8)
    for py in range(2):
        for px in range(4):
            img[py * 16:(py + 1) * 16, px * 16:(px + 1) * 16, :] = py * 4 + px
    pv, grid = synth.patchify(img, mean=(0, 0, 0), std=(1, 1, 1))
    assert grid == (1, 2, 4)
    order = [int(round(float(r[0]) * 255)) for r in pv]
    assert order == [0, 1, 4, 5, 2, 3, 6, 7]
    assert pv.shape == (8, 3 * 2 * 16 * 16)


def test_position_ids_and_mrope_ownership():
    # crane-core/src/models/qwen3_5/vlm.rs:190-241
    ids = [5, 6, 9, 9, 9, 9, 9, 9, 7]
    pos, nxt = ov.build_position_ids(ids, [(1, 4, 6)], 2, 9, 0)
    assert pos[:, :2].tolist() == [[0, 1]] * 3
    assert pos[:, 2:8].tolist() == [[2] * 6, [2, 2, 2, 3, 3, 3], [2, 3, 4, 2, 3, 4]]
    assert nxt == 2 + 3 + 1 and pos[:, 8].tolist() == [5, 5, 5]
    # qwen3_5/modeling.rs:203-233 for mrope_section [11,11,10], half_rot 32
    ax = ov.mrope_axis_of(32, [11, 11, 10])
    assert [i for i, a in enumerate(ax) if a == 1] == list(range(1, 32, 3))
    assert [i for i, a in enumerate(ax) if a == 2] == list(range(2, 30, 3))


# ---- properties the reference's unit tests assert --------------------------------------------------

def test_chunked_prefill_equals_single_and_incremental_decode():
    # qwen3/modeling.rs:1763-1801 (chunked == single, 1e-4); modules/attention.rs:1125-1232 (incremental == prefill, 1e-5)
    cfg = synth.TINY_QWEN3
    w = _weights(cfg)
    ids = synth.synth_token_ids(20, cfg["vocab_size"], "chunk")
    a = oq.Qwen3Oracle(cfg, w)
    full = a.forward(ids, 0).numpy()
    b = oq.Qwen3Oracle(cfg, w)
    b.forward(ids[:7], 0)
    b.forward(ids[7:15], 7)
    part = b.forward(ids[15:], 15).numpy()
    assert rel_err(part, full) < 1e-4
    c = oq.Qwen3Oracle(cfg, w)
    for i, t in enumerate(ids):
        inc = c.forward([t], i).numpy()
    assert rel_err(inc, full) < 1e-4


def test_forward_embeds_equals_forward():
    # qwen3/modeling.rs:1440-1462
    cfg = synth.TINY_QWEN3
    w = _weights(cfg)
    ids = synth.synth_token_ids(9, cfg["vocab_size"], "emb")
    a, b = oq.Qwen3Oracle(cfg, w), oq.Qwen3Oracle(cfg, w)
    assert rel_err(b.forward_embeds(b.embed(ids), 0).numpy(), a.forward(ids, 0).numpy()) < 1e-6


def test_bf16_helpers_roundtrip():
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    bits = synth.f32_to_bf16_bits(x)
    assert np.array_equal(synth.bf16_bits_to_f32(bits), torch.from_numpy(x).to(torch.bfloat16).float().numpy())


# ---- Qwen3.5 hybrid (GDN + gated attention) ---------------------------------------------------------

def test_tiny_qwen3_5_matches_hf_fixture():
    from oracle.qwen3_5 import Qwen3_5Oracle
    cfg = synth.TINY_QWEN3_5
    g = golden("tiny_qwen3_5")
    orc = Qwen3_5Oracle(cfg, _weights(cfg))
    toks = [int(t) for t in g["prompt"]]
    for step in range(g["logits"].shape[0]):
        ctx = toks if step == 0 else toks[-1:]
        lo = orc.forward(ctx, len(toks) - len(ctx)).numpy()
        assert rel_err(lo, g["logits"][step]) < 5e-5
        toks.append(int(g["tokens"][step]))


def test_qwen3_5_chunked_prefill_equals_single_pass():
    # crane-core/src/models/qwen3_5/prefill.rs:146-278 (RandWeights backend, 1e-4): KV + conv + recurrent state hand-off
    from oracle.qwen3_5 import Qwen3_5Oracle
    cfg = synth.TINY_QWEN3_5
    w = _weights(cfg)
    ids = synth.synth_token_ids(23, cfg["vocab_size"], "q35chunk")
    a, b, c = (Qwen3_5Oracle(cfg, w) for _ in range(3))
    full = a.forward(ids, 0).numpy()
    b.forward(ids[:9], 0)
    b.forward(ids[9:10], 9)          # a one-token chunk takes decode_conv1d's path (ops/gdn/conv.rs:82-101)
    part = b.forward(ids[10:], 10).numpy()
    assert rel_err(part, full) < 1e-4
    for i, t in enumerate(ids):
        inc = c.forward([t], i).numpy()
    assert rel_err(inc, full) < 1e-4


def test_gdn_v_head_expansion_order():
    # crane-core/src/ops/gdn/layer.rs:281-293: Interleaved (HF) puts a key head's replicas next to each other
    k = torch.tensor([10.0, 20.0]).view(1, 2, 1)
    assert k.repeat_interleave(2, dim=1).flatten().tolist() == [10, 10, 20, 20]


def test_gdn_recurrence_definition():
    # kernels/cuda/gdn.cu:29-34 / ops/gdn/backend.rs:121-148, one step by hand: S=0, k=e0, v, beta -> S=k (x) beta v ; y = S^T q
    from oracle.qwen3_5 import gated_delta_rule
    K, V = 4, 3
    q = torch.tensor([[[2.0, 0, 0, 0]]]); k = torch.tensor([[[1.0, 0, 0, 0]]]); v = torch.tensor([[[1.0, 2, 3]]])
    y, s = gated_delta_rule(q, k, v, torch.zeros(1, 1), torch.tensor([[0.5]]), torch.zeros(1, K, V))
    assert torch.allclose(s[0, 0], torch.tensor([0.5, 1.0, 1.5])) and torch.allclose(y[0, 0], s[0, 0] * 2.0 / 2.0)


# ---- ggml block formats (A14): pinned byte-for-byte to the `gguf` package's definition ----------------------------

def test_ggml_dequant_matches_gguf_package():
    from gguf import quants, GGMLQuantizationType as T
    from oracle import ggml_quant as gq
    x = (np.random.default_rng(5).standard_normal((5, 1024)) * 0.05).astype(np.float32)
    for name, t, tol in (("Q8_0", T.Q8_0, 0.01), ("Q4_K", T.Q4_K, 0.12), ("Q6_K", T.Q6_K, 0.04)):
        raw = gq.quantize(x, name)
        assert raw.shape == (5, gq.row_bytes(name, 1024))
        mine = gq.dequantize(raw, name, 1024)
        assert np.array_equal(mine, quants.dequantize(raw, t)), name           # layout == ggml's
        assert np.abs(mine - x).max() / np.abs(x).max() < tol, name            # and the encoder is a sane quantiser
    # gguf's own Q8_0 encoder decodes identically through ours
    raw = quants.quantize(x, T.Q8_0)
    assert np.array_equal(gq.dequantize(raw, "Q8_0", 1024), quants.dequantize(raw, T.Q8_0))


# ---- Qwen3-TTS frame loop (oracle-only properties; the decoder stack itself is the HF-pinned Qwen3 oracle) ----------------------

def test_tts_oracle_frame_loop_properties():
    from oracle.qwen3_tts import Qwen3TTSOracle
    cfg = synth.TINY_QWEN3_TTS
    o = Qwen3TTSOracle(cfg, _weights(cfg))
    ids = synth.synth_token_ids(5, 1000, "tts-cpu")
    pre, trail, pad = o.build_prefill_embeds(ids)
    # qwen3_tts/modeling.rs:597-726: role prefix (3) + overlaid codec prefix (nothink, think_bos, think_eos, pad = 4) + first text/bos (1)
    assert pre.shape == (8, 256) and trail.shape == (5, 256) and pad.shape == (256,)
    frames, trace = o.generate_codes(ids, 4, repetition_penalty=1.05)
    G, eos, V = cfg["talker_config"]["num_code_groups"], cfg["talker_config"]["codec_eos_token_id"], cfg["talker_config"]["vocab_size"]
    assert all(len(f) == G for f in frames)
    # suppress window [V-1024, V) \\ {EOS} and EOS held back for the first two frames (:1470-1524)
    for step, t in enumerate(trace):
        lg = t["first_logits"].numpy()
        assert np.all(np.isneginf(lg[[i for i in range(V - 1024, V) if i != eos]]))
        assert (step >= 2) or np.isneginf(lg[eos])
    # teacher forcing reproduces the same logits (KV / code-predictor cache handling is consistent)
    _, trace2 = o.generate_codes(ids, 4, repetition_penalty=1.05, forced_frames=frames)
    assert all(torch.allclose(a["group_logits"], b["group_logits"]) for a, b in zip(trace, trace2))


# ---- sampler oracle against the reference's own literal cases (crane-serve/src/engine/sampling.rs:489-640) -----------------

def test_sampling_penalties_known_answers():
    from oracle import sampling as smp
    assert smp.apply_penalties([1.0, 2.0, 3.0], 1.0, 0.0, 0.0, [0, 1, 2]).tolist() == [1.0, 2.0, 3.0]          # all off (:501-506)
    assert smp.apply_penalties([1.0, 2.0, 3.0], 1.1, 0.5, 0.5, []).tolist() == [1.0, 2.0, 3.0]                  # empty context (:510-514)
    np.testing.assert_allclose(smp.apply_penalties([10.0, -10.0, 3.0], 2.0, 0.0, 0.0, [0, 1]), [5.0, -20.0, 3.0], atol=1e-6)   # :520-536
    np.testing.assert_allclose(smp.apply_penalties([10.0, 10.0], 2.0, 1.0, 0.0, [0, 0, 1]), [3.0, 4.0], atol=1e-6)              # :543-551
    assert smp.apply_penalties([10.0] * 3, 1.0, 0.5, 0.0, [0, 0, 0, 1]).tolist() == [8.5, 9.5, 10.0]             # :557-564
    assert smp.apply_penalties([10.0] * 3, 1.0, 0.0, 0.5, [0, 0, 0, 1]).tolist() == [9.5, 9.5, 10.0]             # :569-576
    out = smp.apply_penalties([5.0, 4.9], 1.0, 0.1, 0.0, [0, 1, 0, 0, 0, 0])                                       # :585-593
    np.testing.assert_allclose(out, [4.5, 4.8], atol=1e-6)
    assert out[1] > out[0]
    np.testing.assert_allclose(smp.apply_penalties([10.0] * 3, 1.0, 0.5, 0.2, [0, 0, 0, 1]), [10 - 1.7, 10 - 0.7, 10.0], atol=1e-6)   # :599-606
    np.testing.assert_allclose(smp.apply_penalties([10.0, 10.0], 1.0, -0.5, 0.0, [0, 0, 1]), [11.0, 10.5], atol=1e-6)          # :612-622
    np.testing.assert_allclose(smp.apply_penalties([10.0] * 3, 1.0, 0.0, -0.5, [0, 1]), [10.5, 10.5, 10.0], atol=1e-6)         # :628-640


def test_sampling_topk_known_answers_and_nucleus_rule():
    from oracle import sampling as smp
    assert smp.topk_indices([0.5, -3.0, 7.25, 1.0, 7.5], 5).tolist() == [4, 2, 3, 0, 1]                           # rocm_kernels.rs:180-189
    v = (np.arange(240_000) % 4).astype(np.float32) * 0.5                                                          # :156-172
    for k in (1, 40, 64):
        assert smp.topk_indices(v, k).tolist() == [j * 4 + 3 for j in range(k)]
    # greedy = lowest index among the maxima; temperature -> Gumbel-max with the supplied uniforms
    assert smp.sample([1.0, 3.0, 3.0, 2.0], 0.0)[0] == 1
    # nucleus: probabilities 0.6439, 0.2369, 0.0871, 0.0321 at T = 1: top_p = 0.7 keeps the first token that crosses p too (sampling.rs:312-330)
    lg = np.array([4.0, 3.0, 2.0, 1.0], np.float32)
    u = np.array([0.5, 0.5, 0.999, 0.999], np.float32)               # huge Gumbel noise on tokens 2 and 3: they only win if they are unmasked
    assert smp.sample(lg, 1.0, top_p=0.7, top_k=4, uniforms=u)[0] in (0, 1)
    assert smp.sample(lg, 1.0, top_p=0.95, top_k=4, uniforms=u)[0] == 2
    assert smp.sample(lg, 1.0, top_p=None, top_k=4, uniforms=u)[0] == 2
