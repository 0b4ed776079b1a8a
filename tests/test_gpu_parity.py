"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the committed fixtures.

Tolerances ("rel" = max|a-b| / max|b| over a logits vector; BASELINE.json's bar is <=1e-3 rel against the f32 CPU path;
the measured values are printed):
  * kernel-level GEMM, f32 output ...... 1e-5  (same bf16 inputs, f32 accumulation, different order)
  * default engine (split precision: hi+lo bf16 operands and KV pages) ... 1e-3 everywhere (measured ~1e-5 .. 3e-5)
  * "precision": "bf16" fast mode (plain bf16 operands / pages) .......... 1e-2 decode, 2e-2 prefill (measured 4e-3 .. 1e-2)
  * GGUF-quantised linears against candle's CPU QMatMul semantics (Q8_K / Q8_0 activation blocks, ggml integer dots) ... 1e-3
Greedy tokens must match the oracle wherever the oracle's own top-2 margin exceeds the measured logit error.
"""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err
import crane_b200
from crane_b200 import synth
from oracle.qwen3 import Qwen3Oracle
from oracle.qwen3_vl import Qwen3VLOracle

pytestmark = pytest.mark.gpu

DECODE_TOL = 1e-3          # default (split-precision) engine: the north-star bar
PREFILL_TOL = 1e-3
FAST_DECODE_TOL = 1e-2     # "precision": "bf16"
FAST_PREFILL_TOL = 2e-2


def _bf16(x):
    return synth.f32_to_bf16_bits(x), synth.bf16_round(x)


@pytest.mark.parametrize("simt", [True, False], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("shape", [(128, 128, 64), (457, 1024, 256), (64, 384, 128), (300, 4096, 2048), (1, 256, 512)])
def test_gemm_store_f32(shape, simt):
    M, N, K = shape
    rng = np.random.default_rng(M * 7 + N)
    a_bits, a = _bf16(rng.standard_normal((M, K), dtype=np.float32))
    w_bits, w = _bf16(rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K))
    bias = rng.standard_normal(N).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    out = crane_b200.op_gemm(a_bits, w_bits, crane_b200.EPI_STORE_F32, use_simt=simt)
    e = rel_err(out, ref)
    print(f"gemm {shape} simt={simt}: rel err {e:.3e}")
    assert e < 1e-5
    out = crane_b200.op_gemm(a_bits, w_bits, crane_b200.EPI_STORE_F32, bias=bias, use_simt=simt)
    assert rel_err(out, ref + bias) < 1e-5


@pytest.mark.parametrize("simt", [True, False], ids=["simt", "tcgen05"])
def test_gemm_split_precision(simt):
    """A = hi + lo (two bf16 planes): the product carries ~16 mantissa bits of the f32 activations."""
    M, N, K = 300, 512, 1024
    rng = np.random.default_rng(11)
    a = rng.standard_normal((M, K), dtype=np.float32)
    hi_bits, hi = _bf16(a)
    lo_bits, lo = _bf16(a - hi)
    w_bits, w = _bf16(rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K))
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    e_plain = rel_err(crane_b200.op_gemm(hi_bits, w_bits, crane_b200.EPI_STORE_F32, use_simt=simt), ref)
    e_split = rel_err(crane_b200.op_gemm(hi_bits, w_bits, crane_b200.EPI_STORE_F32, use_simt=simt, a_lo_bits=lo_bits), ref)
    print(f"gemm split simt={simt}: plain bf16 {e_plain:.3e}, split {e_split:.3e}")
    assert e_split < 3e-5 and e_plain > 10 * e_split


@pytest.mark.parametrize("simt", [True, False], ids=["simt", "tcgen05"])
def test_gemm_epilogues(simt):
    M, N, K = 200, 512, 256
    rng = np.random.default_rng(3)
    a_bits, a = _bf16(rng.standard_normal((M, K), dtype=np.float32))
    w_bits, w = _bf16(rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K))
    bias = rng.standard_normal(N).astype(np.float32)
    ref = (a.astype(np.float64) @ w.astype(np.float64).T)
    init = rng.standard_normal((M, N)).astype(np.float32)
    out = crane_b200.op_gemm(a_bits, w_bits, crane_b200.EPI_RESID_F32, out_init=init, use_simt=simt)
    assert rel_err(out, init + ref) < 1e-5
    out = synth.bf16_bits_to_f32(crane_b200.op_gemm(a_bits, w_bits, crane_b200.EPI_STORE_BF16, use_simt=simt))
    assert rel_err(out, ref) < 5e-3
    g, u = ref[:, 0::2], ref[:, 1::2]
    silu = g / (1 + np.exp(-g)) * u
    out = synth.bf16_bits_to_f32(crane_b200.op_gemm(a_bits, w_bits, crane_b200.EPI_SILU_MUL_BF16, use_simt=simt))
    assert out.shape == (M, N // 2) and rel_err(out, silu) < 5e-3
    x = torch.from_numpy(ref + bias)
    for mode, approx in ((crane_b200.EPI_GELU_ERF_BF16, "none"), (crane_b200.EPI_GELU_TANH_BF16, "tanh")):
        out = synth.bf16_bits_to_f32(crane_b200.op_gemm(a_bits, w_bits, mode, bias=bias, use_simt=simt))
        assert rel_err(out, torch.nn.functional.gelu(x, approximate=approx).numpy()) < 5e-3


@pytest.mark.parametrize("shape", [(454, 2048, 6144), (784, 1024, 4096), (454, 12288, 2048), (300, 1184, 512), (1, 2048, 64), (2000, 4096, 1024)],
                         ids=["down", "vit-fc2", "gate-up", "ragged-n", "one-row-one-kstep", "multi-wave"])
def test_gemm_stream_k_full_width(shape):
    """The persistent stream-K schedule at the widths of the BASELINE models: tiles shared by 2..5 CTAs (partials parked and summed
    in contributor order), CTAs spanning several tiles (double-buffered accumulator), a ragged last column tile, every fused
    epilogue -- against float64, and bitwise equal from run to run."""
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K), dtype=np.float32)
    hi_bits, hi = _bf16(a)
    lo_bits, lo = _bf16(a - hi)
    w_bits, w = _bf16(rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K))
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    init = rng.standard_normal((M, N)).astype(np.float32)
    o1 = crane_b200.op_gemm(hi_bits, w_bits, crane_b200.EPI_RESID_F32, out_init=init, a_lo_bits=lo_bits)
    o2 = crane_b200.op_gemm(hi_bits, w_bits, crane_b200.EPI_RESID_F32, out_init=init, a_lo_bits=lo_bits)
    e = rel_err(o1, init + ref)
    print(f"stream-k {shape}: split resid rel {e:.2e}")
    assert e < 3e-5 and np.array_equal(o1, o2)
    plain = hi.astype(np.float64) @ w.astype(np.float64).T
    assert rel_err(crane_b200.op_gemm(hi_bits, w_bits, crane_b200.EPI_STORE_F32), plain) < 1e-5
    if N % 64 == 0:
        g, u = ref[:, 0::2], ref[:, 1::2]
        out = synth.bf16_bits_to_f32(crane_b200.op_gemm(hi_bits, w_bits, crane_b200.EPI_SILU_MUL_BF16, a_lo_bits=lo_bits))
        assert out.shape == (M, N // 2) and rel_err(out, g / (1 + np.exp(-g)) * u) < 5e-3


# (gemm kernel, precision mode): the default engine twice (debug SIMT GEMM / tcgen05 GEMM) + the plain-bf16 fast mode
MODES = [("simt", "split"), ("tcgen05", "split"), ("tcgen05", "bf16")]
MODE_IDS = ["simt-split", "tcgen05-split", "tcgen05-bf16"]


def _tols(precision):
    return (PREFILL_TOL, DECODE_TOL) if precision == "split" else (FAST_PREFILL_TOL, FAST_DECODE_TOL)


def _model(cfg, cls=crane_b200.Qwen3Model, **opts):
    w = dict(synth.synth_checkpoint(cfg))
    m = cls(cfg, device=0, max_seq_len=opts.pop("max_seq_len", 512), **opts)
    m.load_checkpoint(w.items())
    return m, w


@pytest.mark.parametrize("name,cfg", [("tiny_qwen3", synth.TINY_QWEN3), ("tiny_qwen3_untied", synth.TINY_QWEN3_UNTIED)])
@pytest.mark.parametrize("gemm,precision", MODES, ids=MODE_IDS)
def test_tiny_qwen3_against_hf_fixture(name, cfg, gemm, precision):
    g = golden(name)
    ptol, dtol = _tols(precision)
    m, _ = _model(cfg, gemm=gemm, precision=precision)
    toks = [int(t) for t in g["prompt"]]
    errs = []
    for step in range(g["logits"].shape[0]):
        ctx = toks if step == 0 else toks[-1:]
        lg = m.forward_step(ctx, len(toks) - len(ctx))
        e = rel_err(lg, g["logits"][step])
        errs.append(e)
        top2 = np.sort(g["logits"][step])[-2:]
        if top2[1] - top2[0] > 4 * e * np.abs(g["logits"][step]).max():
            assert int(np.argmax(lg)) == int(g["tokens"][step]), f"step {step}"
        toks.append(int(g["tokens"][step]))
    print(f"{name} gemm={gemm} {precision}: prefill rel {errs[0]:.3e}, decode rel max {max(errs[1:]):.3e}")
    assert errs[0] < ptol and max(errs[1:]) < dtol
    m.close()


def test_chunked_prefill_incremental_decode_and_argmax():
    cfg = synth.TINY_QWEN3
    m, w = _model(cfg)
    orc = Qwen3Oracle(cfg, w)
    ids = synth.synth_token_ids(150, cfg["vocab_size"], "chunk-gpu")     # spans 3 KV pages
    ref = orc.forward(ids, 0).numpy()
    full = m.forward_step(ids, 0)
    assert rel_err(full, ref) < PREFILL_TOL
    m.clear_kv_cache()
    m.forward_step(ids[:70], 0)
    m.forward_step(ids[70:131], 70)
    part = m.forward_step(ids[131:], 131)
    assert rel_err(part, ref) < PREFILL_TOL
    m.clear_kv_cache()
    for i, t in enumerate(ids):                                          # decode kernels only
        inc = m.forward_step([t], i)
    e = rel_err(inc, ref)
    print(f"incremental decode (150 tokens) vs oracle prefill: rel {e:.3e}")
    assert e < DECODE_TOL
    assert m.kv_len() == 150
    m.clear_kv_cache()
    tok = m.forward_step_argmax(ids, 0)
    lg = m.copy_logits()
    assert tok == int(np.flatnonzero(lg == lg.max())[0])                 # lowest index among maxima
    m.close()


def test_long_context_decode_multiple_kv_tiles():
    """Context long enough that every KV split of the decode attention walks several shared-memory tiles."""
    cfg = synth.TINY_QWEN3
    m, w = _model(cfg, max_seq_len=2560)
    orc = Qwen3Oracle(cfg, w, max_pos=2560)
    ids = synth.synth_token_ids(2300, cfg["vocab_size"], "long")
    ref = orc.forward(ids, 0).numpy()
    e0 = rel_err(m.forward_step(ids, 0), ref)
    errs, tok = [], int(np.argmax(ref))
    for i in range(3):
        ref = orc.forward([tok], 2300 + i).numpy()
        errs.append(rel_err(m.forward_step([tok], 2300 + i), ref))
        tok = int(np.argmax(ref))
    print(f"long context (2300 + 3): prefill rel {e0:.3e}, decode rel max {max(errs):.3e}")
    assert e0 < PREFILL_TOL and max(errs) < DECODE_TOL
    m.close()


def test_on_device_greedy_loop_matches_host_loop_and_oracle():
    cfg = synth.TINY_QWEN3
    m, w = _model(cfg)
    orc = Qwen3Oracle(cfg, w)
    ids = synth.synth_token_ids(40, cfg["vocab_size"], "greedy")
    ref_toks, margins = orc.generate_greedy(ids, 24)
    dev = m.generate(ids, max_new_tokens=24)
    # host-driven loop through forward_step_argmax (the server's greedy path)
    m.clear_kv_cache()
    host = [m.forward_step_argmax(ids, 0)]
    for i in range(23):
        host.append(m.forward_step_argmax([host[-1]], len(ids) + i))
    assert list(dev) == host, "on-device loop and host-driven loop disagree"
    # identical to the oracle up to the first near-tie of the oracle itself
    for i, (a, b) in enumerate(zip(dev, ref_toks)):
        if a != b:
            assert margins[i] < 2e-3, f"token {i}: {a} vs oracle {b} with margin {margins[i]}"   # only an oracle near-tie may differ
            break
    print(f"greedy: {sum(int(a == b) for a, b in zip(dev, ref_toks))}/24 tokens equal, min oracle margin {min(margins):.3f}")
    m.close()


@pytest.mark.parametrize("cfg_name", ["tiny_qwen3", "tiny_qwen3_untied"])
def test_persistent_decode_kernel_matches_kernel_chain(cfg_name):
    """The single-launch decode (decode_ll.cu: tagged activation exchange, no grid barrier) against the graph of PDL-chained
    kernels: same tokens from the on-device loop, logits equal to f32 summation-order noise, and both within the bar of the oracle."""
    cfg = synth.TINY_QWEN3 if cfg_name == "tiny_qwen3" else synth.TINY_QWEN3_UNTIED
    w = dict(synth.synth_checkpoint(cfg))
    ids = synth.synth_token_ids(70, cfg["vocab_size"], "ll")
    out = {}
    for persistent in (True, False):
        m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=512, persistent=persistent)
        m.load_checkpoint(w.items())
        first = m.forward_step_argmax(ids, 0)
        toks = list(m.decode_greedy(first, len(ids), 40))                       # 40 steps in one launch (persistent) / 40 graph replays
        lg = m.forward_step([toks[-1]], len(ids) + 40)                           # single step, logits out
        lens = (m.kv_len(), m.kernel_launches())
        more = list(m.decode_greedy(int(np.argmax(lg)), len(ids) + 41, 5))      # a second launch continues from the first one's state
        out[persistent] = (first, toks, lg, more, lens)
        m.close()
    a, b = out[True], out[False]
    e = rel_err(a[2], b[2])
    print(f"persistent vs chain ({cfg_name}): logits rel {e:.3e}; launches {a[4][1]} vs {b[4][1]}")
    # the persistent kernel feeds the tensor cores hi + lo bf16 activations (~16 mantissa bits), the chain multiplies by the f32
    # values: logits agree to the split-precision level, tokens wherever the chain's own top-2 margin is not a near-tie
    assert a[0] == b[0] and a[4][0] == b[4][0] == len(ids) + 41
    same = sum(int(x == y) for x, y in zip(a[1], b[1]))
    print(f"   greedy tokens equal: {same}/40 in the first launch")
    assert a[1][:8] == b[1][:8]
    assert e < 2e-4
    orc = Qwen3Oracle(cfg, w)
    ref = orc.forward(list(ids) + [a[0]] + a[1], 0).numpy()
    assert rel_err(a[2], ref) < DECODE_TOL
    assert a[4][1] < b[4][1] // 10                                               # it really was one launch per call


def test_forward_embeds_and_errors():
    cfg = synth.TINY_QWEN3
    m, w = _model(cfg)
    orc = Qwen3Oracle(cfg, w)
    ids = synth.synth_token_ids(33, cfg["vocab_size"], "embeds")
    ref = orc.forward(ids, 0).numpy()
    emb = orc.embed(ids).numpy()
    assert rel_err(m.forward_embeds(emb, 0), ref) < PREFILL_TOL
    with pytest.raises(crane_b200.CraneB200Error) as e:
        m.forward_step([1, 2], 5)                                        # start_pos must equal the cached length
    assert e.value.code == crane_b200.INVALID_ARG
    with pytest.raises(crane_b200.CraneB200Error):
        m.forward_step([cfg["vocab_size"]], 33)                          # token id out of range
    m.clear_kv_cache()
    assert m.kv_len() == 0 and m.num_layers() == cfg["num_hidden_layers"]
    assert rel_err(m.forward_step(ids, 0), ref) < PREFILL_TOL           # handle still usable after errors
    m.close()


@pytest.mark.parametrize("gemm,precision", MODES, ids=MODE_IDS)
def test_tiny_qwen3_vl_against_fixture(gemm, precision):
    cfg = synth.TINY_QWEN3_VL
    g = golden("tiny_qwen3_vl")
    ptol, dtol = _tols(precision)
    pv, grid = synth.patchify(g["image"])
    m, w = _model(cfg, cls=crane_b200.Qwen3VLModel, gemm=gemm, precision=precision)
    img, ds = m.encode_images(pv, [grid], want_deepstack=3)
    e_img, e_ds = rel_err(img, g["ref_image_embeds"]), rel_err(ds, g["ref_deepstack"])
    lg = m.forward(g["prompt"], pv, [grid], 0)
    e0 = rel_err(lg, g["ref_logits"][0])
    S = len(g["prompt"])
    errs = []
    tok = int(g["ref_tokens"][0])
    for step in range(1, g["ref_logits"].shape[0]):
        lg = m.decode_step(tok, S + step - 1)
        errs.append(rel_err(lg, g["ref_logits"][step]))
        tok = int(g["ref_tokens"][step])
    print(f"tiny_qwen3_vl gemm={gemm} {precision}: image rel {e_img:.3e} deepstack rel {e_ds:.3e} prefill rel {e0:.3e} decode rel {max(errs):.3e}")
    assert e_img < ptol and e_ds < ptol and e0 < ptol and max(errs) < dtol
    # HF's activation choice through the engine switch
    m2, _ = _model(cfg, cls=crane_b200.Qwen3VLModel, gemm=gemm, precision=precision, vit_act="tanh", merger_act="erf")
    assert rel_err(m2.forward(g["prompt"], pv, [grid], 0), g["hf_logits"][0]) < ptol
    m.close(); m2.close()


def test_vl_generate_matches_oracle():
    cfg = synth.TINY_QWEN3_VL
    m, w = _model(cfg, cls=crane_b200.Qwen3VLModel)
    image = synth.synth_image(96, 64, "gen")
    pv, grid = synth.patchify(image)
    ids = synth.build_vl_prompt(cfg, 30, grid, "gen")
    orc = Qwen3VLOracle(cfg, w)
    lo = orc.prefill(ids, pv, [grid]).numpy()
    ref = [int(lo.argmax())]
    for i in range(7):
        ref.append(int(orc.decode_step(ref[-1], len(ids) + i).numpy().argmax()))
    got = m.generate(ids, pv, [grid], 8)
    print("vl generate:", list(got), "oracle:", ref)
    assert [int(x) for x in got] == ref
    # the serving loop's per-token call (4 bytes back per step) walks the same M-RoPE positions as the on-device loop
    m.clear_kv_cache()
    step = [int(np.argmax(m.forward(ids, pv, [grid], 0)))]
    for i in range(7):
        step.append(m.decode_step_argmax(step[-1], len(ids) + i))
    assert step == ref
    m.close()


# ---- Qwen3.5 hybrid: Gated-Delta-Net layers + gated attention with partial rotary (config 3) ----------------

@pytest.mark.parametrize("gemm,precision", MODES, ids=MODE_IDS)
def test_tiny_qwen3_5_against_hf_fixture(gemm, precision):
    cfg = synth.TINY_QWEN3_5
    g = golden("tiny_qwen3_5")
    ptol, _ = _tols(precision)
    m, _ = _model(cfg, cls=crane_b200.Qwen3_5Model, gemm=gemm, precision=precision)
    toks = [int(t) for t in g["prompt"]]
    errs = []
    for step in range(g["logits"].shape[0]):
        ctx = toks if step == 0 else toks[-1:]
        errs.append(rel_err(m.forward_step(ctx, len(toks) - len(ctx)), g["logits"][step]))
        toks.append(int(g["tokens"][step]))
    print(f"tiny_qwen3_5 gemm={gemm} {precision}: prefill rel {errs[0]:.3e}, decode rel max {max(errs[1:]):.3e}")
    assert errs[0] < ptol and max(errs[1:]) < ptol
    m.close()


@pytest.mark.parametrize("equal_heads", [False, True], ids=["nk<nv", "nk==nv"])
def test_qwen3_5_chunked_prefill_state_handoff_and_decode(equal_heads):
    """nk == nv (the shape of Qwen3.5-0.8B) takes the single-launch GDN decode kernel, nk < nv the five-kernel path."""
    from oracle.qwen3_5 import Qwen3_5Oracle
    cfg = dict(synth.TINY_QWEN3_5, linear_num_key_heads=4) if equal_heads else synth.TINY_QWEN3_5
    m, w = _model(cfg, cls=crane_b200.Qwen3_5Model)
    orc = Qwen3_5Oracle(cfg, w)
    ids = synth.synth_token_ids(90, cfg["vocab_size"], "q35-gpu")
    ref = orc.forward(ids, 0).numpy()
    full = m.forward_step(ids, 0)
    m.clear_kv_cache()                                   # must also zero the conv / recurrent state
    m.forward_step(ids[:37], 0)
    m.forward_step(ids[37:38], 37)                       # one-token chunk: decode kernels in the middle of a prefill
    part = m.forward_step(ids[38:], 38)
    m.clear_kv_cache()
    for i, t in enumerate(ids):
        inc = m.forward_step([t], i)
    e = (rel_err(full, ref), rel_err(part, ref), rel_err(inc, ref))
    print(f"qwen3.5 single / chunked / incremental vs oracle: {e[0]:.3e} {e[1]:.3e} {e[2]:.3e}")
    assert max(e) < PREFILL_TOL
    m.clear_kv_cache()
    dev = m.generate(ids, max_new_tokens=12)
    m.clear_kv_cache()
    host = [m.forward_step_argmax(ids, 0)]
    for i in range(11):
        host.append(m.forward_step_argmax([host[-1]], len(ids) + i))
    assert list(dev) == host
    m.close()


@pytest.mark.parametrize("dk", [128, 64, 256])
def test_gdn_chunkwise_recurrence_against_sequential_and_oracle(dk):
    """Prefill of >= 64 rows takes the chunkwise Gated-Delta-Net kernels (gdn_chunk.cu: 64 tokens per serial step on the tensor
    cores); `engine.gdn = "sequential"` keeps the token-by-token kernel.  Both against the oracle, and against each other at the
    reference's own bar for chunked vs sequential (1e-4, crane-core/tests/rocm_kernels.rs:39-84): a 50-row call (sequential in
    either mode) leaves a non-zero state, then 280 rows = 4 chunks + a ragged 24-row tail, then decode steps from the state the
    chunkwise pass left.  Key widths 64 and 256 run the generic-width sequential kernel (the reference accepts K <= 256)."""
    from oracle.qwen3_5 import Qwen3_5Oracle
    cfg = dict(synth.TINY_QWEN3_5, linear_key_head_dim=dk)
    ids = synth.synth_token_ids(330, cfg["vocab_size"], f"gdn-chunk-{dk}")
    outs = {}
    for mode in ("chunked", "sequential"):
        m, w = _model(cfg, cls=crane_b200.Qwen3_5Model, gdn=mode)
        a = m.forward_step(ids[:50], 0)
        b = m.forward_step(ids[50:], 50)
        dec, tok = [], int(np.argmax(b))
        toks = [tok]
        for i in range(4):
            lg = m.forward_step([tok], len(ids) + i)
            dec.append(lg)
            tok = int(np.argmax(lg))
            toks.append(tok)
        outs[mode] = (a, b, dec, toks)
        m.close()
    orc = Qwen3_5Oracle(cfg, w)
    orc.forward(ids[:50], 0)
    ref = orc.forward(ids[50:], 50).numpy()
    e_chunk, e_seq = rel_err(outs["chunked"][1], ref), rel_err(outs["sequential"][1], ref)
    e_cross = rel_err(outs["chunked"][1], outs["sequential"][1])
    e_dec = max(rel_err(x, y) for x, y in zip(outs["chunked"][2], outs["sequential"][2]))
    tok = outs["chunked"][3][0]
    e_dec_orc = []
    for i in range(4):
        r = orc.forward([tok], len(ids) + i).numpy()
        e_dec_orc.append(rel_err(outs["chunked"][2][i], r))
        tok = outs["chunked"][3][i + 1]
    print(f"GDN dk={dk}: chunkwise vs oracle {e_chunk:.2e}, sequential vs oracle {e_seq:.2e}, chunkwise vs sequential {e_cross:.2e}; "
          f"decode after chunkwise prefill vs sequential {e_dec:.2e}, vs oracle {max(e_dec_orc):.2e}")
    assert np.array_equal(outs["chunked"][0], outs["sequential"][0])          # the 50-row call is the same kernel in both modes
    assert e_chunk < PREFILL_TOL and e_seq < PREFILL_TOL and max(e_dec_orc) < DECODE_TOL
    assert e_cross < 1e-4 and e_dec < 1e-4


@pytest.mark.parametrize("nh,nkv", [(6, 2), (12, 2), (16, 2), (10, 2), (24, 4)], ids=["nrep3", "nrep6", "nrep8", "nrep5", "nrep6-24over4"])
def test_decode_attention_any_group_width(nh, nkv):
    """Query groups wider than the attention kernel's sub-group (4 heads) are served by several CTA clusters per KV head: 6 = 2 x 3,
    8 = 2 x 4, 5 = 5 x 1 -- the north star's "4 KV heads broadcast to 24 Q heads" is the last case."""
    cfg = dict(synth.TINY_QWEN3, num_attention_heads=nh, num_key_value_heads=nkv, num_hidden_layers=2)
    m, w = _model(cfg, persistent=False)
    orc = Qwen3Oracle(cfg, w)
    ids = synth.synth_token_ids(70, cfg["vocab_size"], f"nrep{nh}")
    ref = orc.forward(ids, 0).numpy()
    e0 = rel_err(m.forward_step(ids, 0), ref)
    errs, tok = [], int(np.argmax(ref))
    for i in range(5):
        ref = orc.forward([tok], len(ids) + i).numpy()
        errs.append(rel_err(m.forward_step([tok], len(ids) + i), ref))
        tok = int(np.argmax(ref))
    print(f"nh={nh} nkv={nkv}: prefill rel {e0:.2e}, decode rel max {max(errs):.2e}")
    assert e0 < PREFILL_TOL and max(errs) < DECODE_TOL
    m.close()


# ---- full-width geometries of the BASELINE models (few layers, small vocabulary: seconds on the CPU oracle) -----------------------

WIDE_QWEN3_8B = dict(synth.QWEN3_8B, num_hidden_layers=2, vocab_size=4096, max_position_embeddings=4096)
WIDE_QWEN3_5 = dict(synth.QWEN3_5_0_8B, num_hidden_layers=4, vocab_size=4096, max_position_embeddings=4096)
WIDE_QWEN3_VL = dict(synth.QWEN3_VL_2B, image_token_id=4001, vision_start_token_id=4002, vision_end_token_id=4003,
                     text_config=dict(synth.QWEN3_VL_2B["text_config"], num_hidden_layers=2, vocab_size=4096, max_position_embeddings=4096),
                     vision_config=dict(synth.QWEN3_VL_2B["vision_config"], depth=2, deepstack_visual_indexes=[0, 1]))


@pytest.mark.parametrize("persistent", [True, False], ids=["persistent-mlp6144", "chain"])
def test_full_width_qwen3_8b_geometry(persistent):
    """Qwen3-8B's layer at full width (hidden 4096, 32 query heads over 8 KV heads = 4 per group, head_dim 128, MLP 12288), two
    layers: prefill across several KV pages, then decode steps through the kernel chain (attention kernel instance NREP = 4,
    D = 128; the persistent kernel has no 24-slice column split for K = 12288 and declines the shape).  The persistent kernel
    runs the same attention geometry with the MLP at 6144."""
    cfg = dict(WIDE_QWEN3_8B, intermediate_size=6144) if persistent else WIDE_QWEN3_8B
    m, w = _model(cfg)
    assert m.decode_path() == ("persistent" if persistent else "chain")
    orc = Qwen3Oracle(cfg, w)
    ids = synth.synth_token_ids(200, cfg["vocab_size"], "wide8b")
    ref = orc.forward(ids, 0).numpy()
    e0 = rel_err(m.forward_step(ids, 0), ref)
    errs, tok, toks = [], int(np.argmax(ref)), []
    for i in range(6):
        ref = orc.forward([tok], len(ids) + i).numpy()
        errs.append(rel_err(m.forward_step([tok], len(ids) + i), ref))
        tok = int(np.argmax(ref))
        toks.append(tok)
    print(f"Qwen3-8B geometry x2 layers ({'persistent' if persistent else 'chain'}): prefill rel {e0:.2e}, decode rel max {max(errs):.2e}")
    assert e0 < PREFILL_TOL and max(errs) < DECODE_TOL
    m.clear_kv_cache()
    assert [int(t) for t in m.generate(ids, max_new_tokens=7)][1:] == toks
    m.close()


def test_full_width_qwen3_5_geometry():
    """Qwen3.5-0.8B's layers at full width: 16 + 16 linear-attention heads of 128 (nk == nv: the fused GDN decode kernel), gated
    attention with head_dim 256 and 4 query heads per KV head (attention kernel instance NREP = 4, D = 256), partial rotary 0.25."""
    from oracle.qwen3_5 import Qwen3_5Oracle
    cfg = WIDE_QWEN3_5
    m, w = _model(cfg, cls=crane_b200.Qwen3_5Model)
    orc = Qwen3_5Oracle(cfg, w)
    ids = synth.synth_token_ids(150, cfg["vocab_size"], "wide35")
    ref = orc.forward(ids, 0).numpy()
    e0 = rel_err(m.forward_step(ids, 0), ref)
    errs, tok = [], int(np.argmax(ref))
    for i in range(6):
        ref = orc.forward([tok], len(ids) + i).numpy()
        errs.append(rel_err(m.forward_step([tok], len(ids) + i), ref))
        tok = int(np.argmax(ref))
    print(f"Qwen3.5-0.8B geometry x4 layers: prefill rel {e0:.2e}, decode rel max {max(errs):.2e}")
    assert e0 < PREFILL_TOL and max(errs) < DECODE_TOL
    m.close()


def test_full_width_vit_784_patches():
    """The headline request's vision side at full width: one 448 x 448 image = 784 patches of 1536 values through a ViT of width 1024
    (16 heads of 64, MLP 4096, two blocks, both feeding DeepStack mergers), 196 image tokens spliced into a text model of width 2048."""
    cfg = WIDE_QWEN3_VL
    m, w = _model(cfg, cls=crane_b200.Qwen3VLModel)
    image = synth.synth_image(448, 448, "wide-vit")
    pv, grid = synth.patchify(image)
    assert pv.shape[0] == 784
    ids = synth.build_vl_prompt(cfg, 40, grid, "wide-vit")
    orc = Qwen3VLOracle(cfg, w)
    with torch.no_grad():
        ref_img, ref_ds = orc.vision.forward(torch.from_numpy(pv), [grid])
    got_img, got_ds = m.encode_images(pv, [grid], want_deepstack=2)
    e_img = rel_err(got_img, ref_img.numpy())
    e_ds = max(rel_err(g, r.numpy()) for g, r in zip(got_ds, ref_ds))
    ref = orc.prefill(ids, pv, [grid]).numpy()
    e0 = rel_err(m.forward(ids, pv, [grid], 0), ref)
    tok = int(np.argmax(ref))
    e1 = rel_err(m.decode_step(tok, len(ids)), orc.decode_step(tok, len(ids)).numpy())
    print(f"784-patch ViT: image embeddings rel {e_img:.2e}, deepstack rel {e_ds:.2e}; prefill logits rel {e0:.2e}, decode rel {e1:.2e}")
    assert max(e_img, e_ds) < 1e-3 and e0 < PREFILL_TOL and e1 < DECODE_TOL
    m.close()


# ---- GGUF-quantised linears (config 4 path): Q4_K / Q6_K / Q8_0 bytes streamed by the decode GEMV, dequantised for prefill ----

def _quantised_model(cfg, recipe, gemm="tcgen05", **opts):
    """recipe: tensor-name suffix -> ggml type name; returns (engine, oracle weights with dequantised values,
    {name: (raw blocks, type)} for the oracle's integer-dot linears)."""
    from oracle import ggml_quant as gq
    w = dict(synth.synth_checkpoint(cfg))
    m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=512, gemm=gemm, **opts)
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
    return m, wq, qd


Q4_K_M_LIKE = {"q_proj.weight": "Q4_K", "k_proj.weight": "Q4_K", "v_proj.weight": "Q6_K", "o_proj.weight": "Q4_K",
               "gate_proj.weight": "Q4_K", "up_proj.weight": "Q4_K", "down_proj.weight": "Q6_K", "lm_head.weight": "Q6_K"}
ALL_Q8_0 = {k: "Q8_0" for k in Q4_K_M_LIKE} | {"embed_tokens.weight": "Q8_0"}


@pytest.mark.parametrize("qt", ["Q8_0", "Q4_K", "Q6_K"])
@pytest.mark.parametrize("shape", [(5, 256, 96), (4, 4096, 300), (7, 12288, 148 * 2 + 3)], ids=["k256", "k4096-b4", "k12288"])
def test_quantised_linear_kernel_against_integer_dot_oracle(qt, shape):
    """xquant + qgemv on one linear at the widths of the BASELINE models (K = 4096: Qwen3-8B hidden, 12288: its MLP) and with the
    4-row groups of batched decode, against ggml's integer dots (oracle/ggml_quant.py `qmatmul`): identical activation blocks and
    integer sums, so only the order of the f32 additions differs."""
    from oracle import ggml_quant as gq
    m_, k, n = shape
    rng = np.random.default_rng(k + n)
    W = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    x = (rng.standard_normal((m_, k)) * rng.uniform(0.2, 3.0, size=(m_, 1))).astype(np.float32)
    x[0, : k // 2] = 0.0                                            # an all-zero block, and a row whose largest element is negative
    x[-1, 7] = -np.abs(x[-1]).max() * 2
    raw = gq.quantize(W, qt)
    ref = gq.qmatmul(x, raw, qt)
    got = crane_b200.op_qlinear(x, raw, gq.GGML_TYPE_ID[qt], n)
    e = rel_err(got, ref)
    plain = rel_err(x @ gq.dequantize(raw, qt, k).T, ref)
    print(f"qlinear {qt} {shape}: rel {e:.3e} (un-quantised activations would differ by {plain:.1e})")
    assert e < 2e-5
    nw = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    var = (x.astype(np.float32) ** 2).mean(-1, keepdims=True)
    xn = (x * (1.0 / np.sqrt(var + np.float32(1e-6))).astype(np.float32) * nw).astype(np.float32)
    assert rel_err(crane_b200.op_qlinear(x, raw, gq.GGML_TYPE_ID[qt], n, norm_w=nw), gq.qmatmul(xn, raw, qt)) < 1e-3


def _gpu_linear_oracle(cfg, wq, qd):
    """The oracle with every quantised linear ALSO executed by the GPU kernels (xquant + qgemv through crane_b200.op_qlinear) on the
    oracle's own activations; the oracle's result is what flows on, so every linear of the pass is compared on bit-identical inputs
    (`.worst` = largest relative error over the calls, `.calls` = how many)."""
    from oracle import ggml_quant as gq

    class GpuLinearOracle(Qwen3Oracle):
        worst, calls = 0.0, 0

        def _linear(self, full_name, x, w=None):
            y = super()._linear(full_name, x, w)
            if full_name in self.q:
                raw, qt = self.q[full_name]
                g = crane_b200.op_qlinear(x.reshape(-1, x.shape[-1]).numpy(), raw, gq.GGML_TYPE_ID[qt], raw.shape[0])
                self.worst = max(self.worst, rel_err(g, y.reshape(-1, y.shape[-1]).numpy()))
                self.calls += 1
            return y

    return GpuLinearOracle(cfg, wq, quantised=qd)


@pytest.mark.parametrize("recipe_name", ["q4_k_m_like", "all_q8_0"])
def test_quantised_linears_against_integer_dot_oracle(recipe_name):
    """GGUF-quantised models against the oracle that restates candle's CPU `QMatMul` (oracle/ggml_quant.py `qmatmul`: Q8_K / Q8_0
    activation blocks, ggml integer dots).  Three levels, because an int8 activation code is a step function of its input:
      1. every quantised linear of the model, fed the ORACLE's activations, run by the GPU kernels: <= 1e-5 (same blocks, same integer
         sums; only the order of the f32 additions differs) -- this is the parity statement for the arithmetic;
      2. the engine end to end on a 1-layer model, where the linears see inputs that are bit-close to the oracle's: <= 1e-5 -- the
         glue (norm before quantising, 4-row groups, fused epilogues, quantised embedding rows, tied quantised head);
      3. the engine end to end on the 3-layer model: an activation that sits within ~1e-5 (the f32 / split-bf16 noise between two
         correct implementations) of a rounding boundary takes the neighbouring int8 code, and one flipped code moves an output
         by ~1e-3; a few of them per pass leave the logits within the scale of the quantisation noise itself, which is what is
         asserted, with the distance to the un-quantised-activation semantics (round 1's oracle) printed beside it."""
    cfg = synth.TINY_QWEN3_UNTIED if recipe_name == "q4_k_m_like" else synth.TINY_QWEN3
    recipe = Q4_K_M_LIKE if recipe_name == "q4_k_m_like" else ALL_Q8_0
    m, wq, qd = _quantised_model(cfg, recipe)
    orc = Qwen3Oracle(cfg, wq, quantised=qd)
    plain = Qwen3Oracle(cfg, wq)                                  # y = x_f32 . dequant(W)^T: what round 1 compared against
    ids = synth.synth_token_ids(70, cfg["vocab_size"], "quant")
    ref = orc.forward(ids, 0).numpy()
    # 1. kernels on the oracle's activations
    chk = _gpu_linear_oracle(cfg, wq, qd)
    chk.forward(ids, 0)
    e_kernels = chk.worst
    assert chk.calls >= 7 * cfg["num_hidden_layers"] and e_kernels < 1e-5, (chk.calls, e_kernels)
    # 3. engine, full depth
    gap = rel_err(plain.forward(ids, 0).numpy(), ref)
    e_pre = rel_err(m.forward_step(ids, 0), ref)
    tok = int(ref.argmax())
    errs = []
    for i in range(6):
        ref = orc.forward([tok], len(ids) + i).numpy()
        errs.append(rel_err(m.forward_step([tok], len(ids) + i), ref))
        tok = int(ref.argmax())
    print(f"quantised {recipe_name}: kernels on oracle activations {e_kernels:.1e}; engine prefill rel {e_pre:.3e}, decode rel max {max(errs):.3e} "
          f"(activation quantisation moves the logits by {gap:.1e})")
    assert e_pre < 1.5 * gap and max(errs) < 1.5 * gap
    m.clear_kv_cache()
    dev = m.generate(ids, max_new_tokens=8)
    m.clear_kv_cache()
    host = [m.forward_step_argmax(ids, 0)]
    for i in range(7):
        host.append(m.forward_step_argmax([host[-1]], len(ids) + i))
    assert list(dev) == host
    m.close()
    # 2. engine glue, one layer: a single 4-row group and the single-row kernels
    c1 = dict(cfg, num_hidden_layers=1)
    m1, wq1, qd1 = _quantised_model(c1, recipe)
    o1 = Qwen3Oracle(c1, wq1, quantised=qd1)
    for S in (1, 4):
        e1 = []
        for seed in range(4):                                     # a prompt whose activations avoid every rounding boundary is exact
            ids1 = synth.synth_token_ids(S, c1["vocab_size"], f"qdbg{seed}")
            o1.clear_kv_cache(); m1.clear_kv_cache()
            e1.append(rel_err(m1.forward_step(ids1, 0), o1.forward(ids1, 0).numpy()))
        print(f"   1-layer engine, S={S}: rel " + " ".join(f"{e:.1e}" for e in e1))
        assert sum(e < 1e-5 for e in e1) >= 2 and max(e1) < 1.5 * gap
    m1.close()


def test_gguf_tensor_names_are_accepted():
    from oracle import ggml_quant as gq
    cfg = synth.TINY_QWEN3_UNTIED
    w = dict(synth.synth_checkpoint(cfg))
    ren = {"self_attn.q_proj": "attn_q", "self_attn.k_proj": "attn_k", "self_attn.v_proj": "attn_v", "self_attn.o_proj": "attn_output",
           "self_attn.q_norm": "attn_q_norm", "self_attn.k_norm": "attn_k_norm", "mlp.gate_proj": "ffn_gate", "mlp.up_proj": "ffn_up",
           "mlp.down_proj": "ffn_down", "input_layernorm": "attn_norm", "post_attention_layernorm": "ffn_norm"}
    m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=256)
    for name, arr in w.items():
        g = {"model.embed_tokens.weight": "token_embd.weight", "model.norm.weight": "output_norm.weight", "lm_head.weight": "output.weight"}.get(name)
        if g is None:
            parts = name.split(".")                               # model.layers.N.<...>.weight
            key = ".".join(parts[3:-1])
            g = f"blk.{parts[2]}.{ren[key]}.weight"
        if arr.ndim == 2 and "embd" not in g:
            m.load_tensor_ggml(g, gq.GGML_TYPE_ID["Q8_0"], arr.shape, gq.quantize(arr, "Q8_0"))
        else:
            m.load_tensor(g, arr)
    m.finalize()
    assert len(m.generate(synth.synth_token_ids(12, cfg["vocab_size"], "gguf"), max_new_tokens=4)) == 4
    m.close()


# ---- sequence slots + batched decode (A13: setup/step/extract_batch_decode without padding or KV copies) ------------

@pytest.mark.parametrize("quant", [False, True], ids=["bf16", "q8_0"])
def test_batched_decode_matches_per_sequence_decode(quant):
    cfg = synth.TINY_QWEN3
    if quant:
        m, w, _ = _quantised_model(cfg, ALL_Q8_0)
        m.close()
        from oracle import ggml_quant as gq
        wsrc = dict(synth.synth_checkpoint(cfg))
        m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=256, max_batch=8)
        for name, arr in wsrc.items():
            if arr.ndim == 2:
                m.load_tensor_ggml(name, gq.GGML_TYPE_ID["Q8_0"], arr.shape, gq.quantize(arr, "Q8_0"))
            else:
                m.load_tensor(name, arr)
        m.finalize()
    else:
        w = dict(synth.synth_checkpoint(cfg))
        m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=256, max_batch=8)
        m.load_checkpoint(w.items())
    n, steps = 7, 6                                               # 7 sequences -> groups of 4 + 2 + 1, ragged prompt lengths
    prompts = [synth.synth_token_ids(5 + 9 * i, cfg["vocab_size"], f"b{i}") for i in range(n)]
    # reference: each sequence alone through the single-sequence path
    solo = []
    for p in prompts:
        m.clear_kv_cache()
        solo.append(list(m.generate(p, max_new_tokens=steps + 1)))
    m.clear_kv_cache()
    seqs, first = [], []
    for p in prompts:
        s = m.seq_create()
        m.seq_select(s)
        first.append(m.forward_step_argmax(p, 0))                 # prefill each sequence into its own pages
        seqs.append(s)
    toks, logits = m.decode_batch(seqs, first, n_steps=steps, want_logits=True)
    for i in range(n):
        assert [first[i]] + list(toks[i]) == solo[i], f"sequence {i}"
    assert logits.shape == (n, cfg["vocab_size"]) and np.isfinite(logits).all()
    # a freed slot is reusable and sequence 0 still works
    m.seq_free(seqs[-1])
    s = m.seq_create()
    assert s == seqs[-1]
    m.seq_select(0)
    assert m.kv_len() == 0 and len(m.generate(prompts[0], max_new_tokens=3)) == 3
    with pytest.raises(crane_b200.CraneB200Error):
        m.decode_batch([seqs[0], seqs[0]], [1, 2])
    m.close()


# ---- Qwen3-TTS codec-LM frame loop (config 5): talker + 16-pass code predictor on the device ----------------------------------

@pytest.mark.parametrize("bits", [8, 4], ids=["int8", "int4"])
@pytest.mark.parametrize("kind", ["dense", "hybrid"])
def test_quantised_kv_pages_against_oracle(kind, bits):
    """engine.kv_cache = int8 / int4 = QuantKvCache (qwen3_5/kv_cache.rs:209-342): per (token, head) codes + f32 scale, attention
    reads code * scale.  Single-pass prefill, chunked prefill (the prefix is dequantised from its pages), decode steps (codes are
    dequantised while the attention kernel stages its tile; the new row is quantised in the kernel), fork and export / import.
    A code is a step function of its input: the K/V rows of two correct implementations differ by ~1e-5 (f32 summation order, the
    split-bf16 GEMM operands), a value that close to a rounding boundary takes the neighbouring code (one int8 step is 0.8 % of the
    row's largest element, one int4 step 14 %), and with ~1e-3 of the elements affected the logits move by a few percent of what
    the quantisation itself moves them.  Asserted: the error stays below a quarter of the quantisation's own effect (printed) and
    below 5e-3; the arithmetic (scale, rounding, packing, dequantisation) is exact -- export / import reproduces itself bitwise."""
    if kind == "dense":
        cfg, cls, orc_cls = synth.TINY_QWEN3, crane_b200.Qwen3Model, Qwen3Oracle
    else:
        from oracle.qwen3_5 import Qwen3_5Oracle as orc_cls
        cfg, cls = synth.TINY_QWEN3_5, crane_b200.Qwen3_5Model
    m, w = _model(cfg, cls=cls, kv_cache=f"int{bits}", max_batch=2)
    assert m.decode_path() == "chain"
    orc = orc_cls(cfg, w, kv_bits=bits)
    lossless = orc_cls(cfg, w)
    ids = synth.synth_token_ids(150, cfg["vocab_size"], "kvq")
    ref = orc.forward(ids, 0).numpy()
    gap = rel_err(lossless.forward(ids, 0).numpy(), ref)
    tol = min(5e-3, 0.25 * gap)
    e_full = rel_err(m.forward_step(ids, 0), ref)
    m.clear_kv_cache()
    m.forward_step(ids[:70], 0)
    m.forward_step(ids[70:71], 70)
    e_chunk = rel_err(m.forward_step(ids[71:], 71), ref)
    errs, tok = [], int(np.argmax(ref))
    f = m.seq_fork(0)
    for i in range(5):
        ref = orc.forward([tok], len(ids) + i).numpy()
        errs.append(rel_err(m.forward_step([tok], len(ids) + i), ref))
        if i == 1:                                       # the fork must see the same quantised prefix
            m.seq_select(f)
            o2 = orc_cls(cfg, w, kv_bits=bits)
            o2.forward(ids, 0)
            errs.append(rel_err(m.forward_step([7], len(ids)), o2.forward([7], len(ids)).numpy()))
            m.seq_select(0)
        tok = int(np.argmax(ref))
    caches = m.get_kv_caches()
    m2, _ = _model(cfg, cls=cls, kv_cache=f"int{bits}")
    m2.set_kv_caches(caches, m.kv_len())
    nxt = orc.forward([tok], m.kv_len()).numpy()
    e_imp = rel_err(m2.forward_step([tok], m.kv_len()), nxt)
    print(f"kv int{bits} {kind}: prefill rel {e_full:.2e}, chunked {e_chunk:.2e}, decode / fork max {max(errs):.2e}, after export+import {e_imp:.2e} "
          f"(the quantisation itself moves the logits by {gap:.1e}); cache bytes {m.active_kv_cache_bytes()}")
    assert max(e_full, e_chunk, max(errs), e_imp) < tol
    m.close(); m2.close()


@pytest.mark.parametrize("kind", ["dense", "hybrid"])
def test_seq_fork_shares_a_prefix_and_then_diverges(kind):
    """crane_b200_seq_fork: the fork continues exactly like its source (KV pages; for the hybrid model also the parked Gated-Delta-Net
    conv / recurrent state of every slot), and afterwards the two sequences do not see each other."""
    if kind == "dense":
        cfg, cls = synth.TINY_QWEN3, crane_b200.Qwen3Model
        orc_cls = Qwen3Oracle
    else:
        from oracle.qwen3_5 import Qwen3_5Oracle as orc_cls
        cfg, cls = synth.TINY_QWEN3_5, crane_b200.Qwen3_5Model
    m, w = _model(cfg, cls=cls, max_batch=3)
    orc = orc_cls(cfg, w)
    ids = synth.synth_token_ids(75, cfg["vocab_size"], "fork")       # the prefix ends inside its second KV page
    orc.forward(ids, 0)
    m.forward_step(ids, 0)
    f = m.seq_fork(0)
    assert f == 1
    a_tok, b_tok = [5, 17, 200], [9, 300, 4]
    m.seq_select(f)
    assert m.kv_len() == len(ids)
    got_b = [m.forward_step([t], len(ids) + i).copy() for i, t in enumerate(b_tok)]
    m.seq_select(0)
    got_a = [m.forward_step([t], len(ids) + i).copy() for i, t in enumerate(a_tok)]
    g = m.seq_fork(f)                                                 # fork of a fork, taken while another slot is current
    m.seq_select(g)
    got_b2 = m.forward_step([77], len(ids) + 3).copy()
    ref_a = [orc.forward([t], len(ids) + i).numpy() for i, t in enumerate(a_tok)]
    orc2 = orc_cls(cfg, w)
    orc2.forward(ids, 0)
    ref_b = [orc2.forward([t], len(ids) + i).numpy() for i, t in enumerate(b_tok)]
    ref_b2 = orc2.forward([77], len(ids) + 3).numpy()
    ea = max(rel_err(x, r) for x, r in zip(got_a, ref_a))
    eb = max(rel_err(x, r) for x, r in zip(got_b, ref_b))
    print(f"seq_fork {kind}: source rel {ea:.2e}, fork rel {eb:.2e}, fork-of-fork rel {rel_err(got_b2, ref_b2):.2e}")
    assert max(ea, eb, rel_err(got_b2, ref_b2)) < DECODE_TOL
    m.seq_free(f); m.seq_free(g)
    with pytest.raises(crane_b200.CraneB200Error):
        m.seq_fork(2)                                                 # not a live sequence
    m.close()


@pytest.mark.parametrize("kind", ["dense", "hybrid"])
def test_kv_swap_export_import_round_trip(kind):
    """get_kv_caches / set_kv_caches (backend.rs:65-84): per-layer [n_kv, T, D] tensors (GDN layers: conv window + recurrent state);
    a fresh handle that imports them continues the sequence exactly like the one that computed them."""
    if kind == "dense":
        cfg, cls, orc_cls = synth.TINY_QWEN3, crane_b200.Qwen3Model, Qwen3Oracle
    else:
        from oracle.qwen3_5 import Qwen3_5Oracle as orc_cls
        cfg, cls = synth.TINY_QWEN3_5, crane_b200.Qwen3_5Model
    m, w = _model(cfg, cls=cls)
    ids = synth.synth_token_ids(70, cfg["vocab_size"], "swap")
    m.forward_step(ids, 0)
    caches = m.get_kv_caches()
    assert len(caches) == cfg["num_hidden_layers"]
    full = [i for i in range(len(caches)) if synth.is_full_attention_layer(cfg, i)] if kind == "hybrid" else list(range(len(caches)))
    for i in full:
        assert caches[i][0].shape == (cfg["num_key_value_heads"], 70, cfg["head_dim"])
    ref = [m.forward_step([t], 70 + i).copy() for i, t in enumerate([3, 99, 512])]
    m2, _ = _model(cfg, cls=cls)
    m2.set_kv_caches(caches, 70)
    assert m2.kv_len() == 70
    got = [m2.forward_step([t], 70 + i).copy() for i, t in enumerate([3, 99, 512])]
    e = max(rel_err(g, r) for g, r in zip(got, ref))
    print(f"kv swap {kind}: continuation after import rel {e:.2e}")
    assert e < 1e-6                                       # f32 export of hi + lo re-splits into the same two planes
    m.close(); m2.close()


def test_tts_frame_loop_against_oracle():
    from oracle.qwen3_tts import Qwen3TTSOracle
    cfg = synth.TINY_QWEN3_TTS
    w = dict(synth.synth_checkpoint(cfg))
    m = crane_b200.Qwen3TTSModel(cfg, device=0, max_seq_len=256)
    m.load_checkpoint(w.items())
    orc = Qwen3TTSOracle(cfg, w)
    ids = synth.synth_token_ids(7, cfg["talker_config"]["text_vocab_size"] - 8, "tts-gpu")
    # host glue + text projection / embedding gathers
    pre_o, trail_o, pad_o = orc.build_prefill_embeds(ids)
    pre, trail, pad = m.build_prefill_embeds(list(ids))
    e_glue = max(rel_err(pre, pre_o.numpy()), rel_err(trail, trail_o.numpy()), rel_err(pad, pad_o.numpy()))
    # greedy oracle frames, then teacher-force the same codes through the engine and compare every head's logits
    n = 12
    frames_o, trace = orc.generate_codes(ids, n, repetition_penalty=1.05)
    assert len(frames_o) == n
    frames, fl, gl = m.generate_codes(ids, n, repetition_penalty=1.05, forced_frames=frames_o, want_logits=True)
    assert frames.shape == (n, cfg["talker_config"]["num_code_groups"]) and np.array_equal(frames, np.array(frames_o, np.uint32))
    raw_first = [(t["hidden"] @ orc.w["talker.codec_head.weight"].T).numpy() for t in trace]
    e_first = max(rel_err(fl[i], raw_first[i]) for i in range(n))
    e_group = max(rel_err(gl[i], trace[i]["group_logits"].numpy()) for i in range(n))
    print(f"tts: glue {e_glue:.3e}, first-code logits {e_first:.3e}, code-predictor logits {e_group:.3e}")
    assert e_glue < PREFILL_TOL and e_first < PREFILL_TOL and e_group < PREFILL_TOL
    # free-running greedy on the device: equal to the oracle's greedy frames at least up to the first NEAR-TIE -- the first frame in
    # which some head's top-2 gap (first code: after suppress mask / EOS rule / repetition penalty; the 15 predictor heads: raw) is
    # not clearly above the logit error just measured.  Past such a frame both continuations are legitimate greedy decodes.
    def gap(lg):
        lg = np.asarray(lg, np.float64).reshape(-1)
        fin = lg[np.isfinite(lg)]
        top2 = np.partition(fin, -2)[-2:]
        return float(top2[1] - top2[0]) / float(np.abs(fin).max())
    err = max(e_first, e_group)
    safe = 0
    for t in trace:
        if min([gap(t["first_logits"].numpy())] + [gap(row) for row in t["group_logits"].numpy()]) < 20 * err:
            break
        safe += 1
    g = m.generate_codes(ids, n, repetition_penalty=1.05)
    same = 0
    for a, b in zip(g.tolist(), frames_o):
        if a != b:
            break
        same += 1
    print(f"tts greedy: {same}/{n} frames identical to the oracle; the oracle's first near-tie (top-2 gap < 20 x {err:.1e}) is in frame {safe}")
    assert same >= safe and same >= 1 and g.shape[1] == cfg["talker_config"]["num_code_groups"]
    m.close()


# ---- native checkpoint readers: a .safetensors file and a GGUF file read by the library itself ---------------------

HF_TO_GGUF = {"input_layernorm.weight": "attn_norm.weight", "post_attention_layernorm.weight": "ffn_norm.weight",
              "self_attn.q_proj.weight": "attn_q.weight", "self_attn.k_proj.weight": "attn_k.weight", "self_attn.v_proj.weight": "attn_v.weight",
              "self_attn.o_proj.weight": "attn_output.weight", "self_attn.q_norm.weight": "attn_q_norm.weight",
              "self_attn.k_norm.weight": "attn_k_norm.weight", "mlp.gate_proj.weight": "ffn_gate.weight", "mlp.up_proj.weight": "ffn_up.weight",
              "mlp.down_proj.weight": "ffn_down.weight"}


def _gguf_name(hf):
    if hf == "model.embed_tokens.weight":
        return "token_embd.weight"
    if hf == "model.norm.weight":
        return "output_norm.weight"
    if hf == "lm_head.weight":
        return "output.weight"
    _, _, idx, rest = hf.split(".", 3)
    return f"blk.{idx}.{HF_TO_GGUF[rest]}"


def test_safetensors_file_reader(tmp_path):
    import safetensors.torch
    cfg = synth.TINY_QWEN3_UNTIED
    w = dict(synth.synth_checkpoint(cfg))
    ids = synth.synth_token_ids(33, cfg["vocab_size"], "st")
    ref_m, _ = _model(cfg)
    ref = ref_m.forward_step(ids, 0)
    ref_m.close()
    # bf16 matrices, f32 vectors, plus a tensor the engine has no use for
    tensors = {k: (torch.from_numpy(v).to(torch.bfloat16) if v.ndim == 2 else torch.from_numpy(v)) for k, v in w.items()}
    tensors["model.rotary_emb.inv_freq"] = torch.arange(64, dtype=torch.float32)
    path = str(tmp_path / "model.safetensors")
    safetensors.torch.save_file(tensors, path)
    m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=512)
    loaded, skipped = m.load_safetensors(path)
    m.finalize()
    assert loaded == len(w) and skipped == 1
    assert np.array_equal(m.forward_step(ids, 0), ref)           # same bytes in, same logits out
    with pytest.raises(crane_b200.CraneB200Error):
        m.load_safetensors(str(tmp_path / "missing.safetensors"))
    m.close()


def test_gguf_file_reader(tmp_path):
    import gguf
    from oracle import ggml_quant as gq
    cfg = synth.TINY_QWEN3_UNTIED
    w = dict(synth.synth_checkpoint(cfg))
    ids = synth.synth_token_ids(33, cfg["vocab_size"], "gg")
    ref_m, _, _ = _quantised_model(cfg, Q4_K_M_LIKE)              # the same blocks, registered tensor by tensor
    ref = ref_m.forward_step(ids, 0)
    ref_tok = ref_m.generate(ids, max_new_tokens=4)
    ref_m.close()
    path = str(tmp_path / "model.gguf")
    wr = gguf.GGUFWriter(path, "qwen3")
    wr.add_uint32("qwen3.block_count", cfg["num_hidden_layers"])
    wr.add_uint32("qwen3.embedding_length", cfg["hidden_size"])
    wr.add_uint32("qwen3.feed_forward_length", cfg["intermediate_size"])
    wr.add_uint32("qwen3.attention.head_count", cfg["num_attention_heads"])
    wr.add_uint32("qwen3.attention.head_count_kv", cfg["num_key_value_heads"])
    wr.add_uint32("qwen3.attention.key_length", cfg["head_dim"])
    wr.add_float32("qwen3.rope.freq_base", float(cfg["rope_theta"]))
    wr.add_float32("qwen3.attention.layer_norm_rms_epsilon", float(cfg["rms_norm_eps"]))
    wr.add_string("general.name", "tiny")
    wr.add_array("tokenizer.ggml.tokens", ["a", "b", "c"])        # metadata of every value kind has to be skipped correctly
    qtypes = {"Q4_K": gguf.GGMLQuantizationType.Q4_K, "Q6_K": gguf.GGMLQuantizationType.Q6_K, "Q8_0": gguf.GGMLQuantizationType.Q8_0}
    for name, arr in w.items():
        qt = next((t for suf, t in Q4_K_M_LIKE.items() if name.endswith(suf)), None)
        if qt is None or arr.ndim != 2:
            wr.add_tensor(_gguf_name(name), np.ascontiguousarray(arr, dtype=np.float32))
        else:
            raw = np.ascontiguousarray(gq.quantize(arr, qt)).reshape(arr.shape[0], -1)
            wr.add_tensor(_gguf_name(name), raw, raw_dtype=qtypes[qt])
    wr.add_tensor("rope_freqs.weight", np.ones(64, np.float32))
    wr.write_header_to_file()
    wr.write_kv_data_to_file()
    wr.write_tensors_to_file()
    wr.close()
    m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=512)
    loaded, skipped = m.load_gguf(path)
    m.finalize()
    assert loaded == len(w) and skipped == 1
    assert np.array_equal(m.forward_step(ids, 0), ref)
    m.clear_kv_cache()
    assert list(m.generate(ids, max_new_tokens=4)) == list(ref_tok)
    m.close()
    # config from the file's own metadata (Qwen3Model::from_gguf): no config.json involved
    m2 = crane_b200.Qwen3Model.from_gguf(path, device=0, max_seq_len=512)
    assert rel_err(m2.forward_step(ids, 0), ref) < 1e-6          # (f32 epsilon / rope base read back from the metadata)
    m2.close()


def test_qwen3_5_gguf_file(tmp_path):
    """A llama.cpp `qwen35` GGUF (Qwen3_5Model::from_gguf, qwen3_5/model.rs:155-325, modeling.rs:379-411,688-775): config from the
    metadata alone (layer layout from `ssm_a` presence, mrope sections from an array value), llama.cpp tensor names, norms with the
    +1 already folded in, `ssm_a = -exp(A_log)`, a 2-D conv kernel, and the value heads in `Chunked` order (nk = 2 < nv = 4 here,
    so the order matters) -- must reproduce the safetensors-named model; a Q8_0 / Q4_K / Q6_K variant of the linears
    is dequantised (and rounded to bf16) at load and must match the oracle run on those weights."""
    import gguf
    from oracle import ggml_quant as gq
    from oracle.qwen3_5 import Qwen3_5Oracle
    cfg = synth.TINY_QWEN3_5
    w = dict(synth.synth_checkpoint(cfg))
    ids = synth.synth_token_ids(40, cfg["vocab_size"], "gg35")
    ref_m, _ = _model(cfg, cls=crane_b200.Qwen3_5Model)
    ref = ref_m.forward_step(ids, 0).copy()
    ref_m.close()
    nk, nv, dk, dv = cfg["linear_num_key_heads"], cfg["linear_num_value_heads"], cfg["linear_key_head_dim"], cfg["linear_value_head_dim"]
    vpg, key2 = nv // nk, 2 * nk * dk
    order = [kh * vpg + r for r in range(vpg) for kh in range(nk)]        # chunked slot (r * nk + kh) holds HF head kh * vpg + r

    def chunk_rows(a, first, blk):                                        # HF (interleaved) -> llama.cpp (chunked) over row blocks
        a = a.copy()
        v = a[first:first + nv * blk].reshape(nv, blk, *a.shape[1:])
        a[first:first + nv * blk] = v[order].reshape(nv * blk, *a.shape[1:])
        return a

    ren = {"self_attn.q_proj": "attn_q", "self_attn.k_proj": "attn_k", "self_attn.v_proj": "attn_v", "self_attn.o_proj": "attn_output",
           "self_attn.q_norm": "attn_q_norm", "self_attn.k_norm": "attn_k_norm", "mlp.gate_proj": "ffn_gate", "mlp.up_proj": "ffn_up",
           "mlp.down_proj": "ffn_down", "input_layernorm": "attn_norm", "post_attention_layernorm": "post_attention_norm",
           "linear_attn.in_proj_qkv": "attn_qkv", "linear_attn.in_proj_z": "attn_gate", "linear_attn.in_proj_b": "ssm_beta",
           "linear_attn.in_proj_a": "ssm_alpha", "linear_attn.norm": "ssm_norm", "linear_attn.out_proj": "ssm_out"}

    def convert(name, a):
        a = np.asarray(a, dtype=np.float32)
        if name == "model.embed_tokens.weight":
            return "token_embd.weight", a
        if name == "model.norm.weight":
            return "output_norm.weight", a + 1.0
        i, t = name.split(".")[2], ".".join(name.split(".")[3:])
        if t == "linear_attn.conv1d.weight":
            return f"blk.{i}.ssm_conv1d.weight", chunk_rows(a.reshape(a.shape[0], -1), key2, dv)
        if t == "linear_attn.dt_bias":
            return f"blk.{i}.ssm_dt.bias", chunk_rows(a, 0, 1)
        if t == "linear_attn.A_log":
            return f"blk.{i}.ssm_a", chunk_rows(-np.exp(a), 0, 1)
        stem, suffix = t.rsplit(".", 1)
        if stem in ("input_layernorm", "post_attention_layernorm", "self_attn.q_norm", "self_attn.k_norm"):
            a = a + 1.0                                                   # folded
        elif stem == "linear_attn.in_proj_qkv":
            a = chunk_rows(a, key2, dv)
        elif stem == "linear_attn.in_proj_z":
            a = chunk_rows(a, 0, dv)
        elif stem in ("linear_attn.in_proj_b", "linear_attn.in_proj_a"):
            a = chunk_rows(a, 0, 1)
        elif stem == "linear_attn.out_proj":
            a = np.ascontiguousarray(chunk_rows(np.ascontiguousarray(a.T), 0, dv).T)
        return f"blk.{i}.{ren[stem]}.{suffix}", a

    def write(path, quant):
        wr = gguf.GGUFWriter(path, "qwen35")
        rp = cfg["rope_parameters"]
        for k, v in (("block_count", cfg["num_hidden_layers"]), ("embedding_length", cfg["hidden_size"]), ("feed_forward_length", cfg["intermediate_size"]),
                     ("attention.head_count", cfg["num_attention_heads"]), ("attention.head_count_kv", cfg["num_key_value_heads"]),
                     ("attention.key_length", cfg["head_dim"]), ("rope.dimension_count", int(cfg["head_dim"] * rp["partial_rotary_factor"])),
                     ("ssm.conv_kernel", cfg["linear_conv_kernel_dim"]), ("ssm.state_size", dk), ("ssm.group_count", nk), ("ssm.time_step_rank", nv),
                     ("ssm.inner_size", nv * dv), ("full_attention_interval", cfg["full_attention_interval"]), ("context_length", 4096)):
            wr.add_uint32("qwen35." + k, v)
        wr.add_float32("qwen35.rope.freq_base", float(rp["rope_theta"]))
        wr.add_float32("qwen35.attention.layer_norm_rms_epsilon", float(cfg["rms_norm_eps"]))
        wr.add_array("qwen35.rope.dimension_sections", [int(x) for x in rp["mrope_section"]] + [0])
        wr.add_array("tokenizer.ggml.tokens", ["a", "b"])
        deq = {}
        qtypes = {"Q4_K": gguf.GGMLQuantizationType.Q4_K, "Q6_K": gguf.GGMLQuantizationType.Q6_K, "Q8_0": gguf.GGMLQuantizationType.Q8_0}
        for name, arr in w.items():
            gname, a = convert(name, arr)
            qt = None
            if quant and a.ndim == 2 and a.shape[1] % 256 == 0 and not gname.endswith(("ssm_beta.weight", "ssm_alpha.weight", "ssm_conv1d.weight")):
                qt = "Q6_K" if "attn_v" in gname or "ffn_down" in gname else "Q8_0" if "ssm_out" in gname else "Q4_K"
            if qt:
                raw = np.ascontiguousarray(gq.quantize(a, qt)).reshape(a.shape[0], -1)
                wr.add_tensor(gname, raw, raw_dtype=qtypes[qt])
                deq[name] = qt
            else:
                wr.add_tensor(gname, np.ascontiguousarray(a))
        wr.write_header_to_file(); wr.write_kv_data_to_file(); wr.write_tensors_to_file(); wr.close()
        return deq

    p32 = str(tmp_path / "q35_f32.gguf")
    write(p32, False)
    got_cfg = crane_b200.gguf_config(p32)
    assert got_cfg["model_type"] == "qwen3_5_text" and got_cfg["layer_types"] == ["linear_attention"] * 3 + ["full_attention"]
    assert got_cfg["linear_num_value_heads"] == nv and got_cfg["linear_value_head_dim"] == dv and got_cfg["rope_parameters"]["mrope_section"][:3] == [11, 11, 10]
    m = crane_b200.Qwen3_5Model.from_gguf(p32, device=0, max_seq_len=512)
    got = m.forward_step(ids, 0)
    e = rel_err(got, ref)
    print(f"qwen35 gguf (f32 tensors, chunked v-heads, folded norms): rel {e:.1e} against the HF-named model")
    assert e < 1e-4                                       # (numpy's exp / +1 and the loader's differ in the last bit)
    m.close()
    # quantised linears: dequantised at load
    pq = str(tmp_path / "q35_q.gguf")
    deq = write(pq, True)
    assert len(deq) >= 20
    wq = dict(w)
    for name, qt in deq.items():
        gname, a = convert(name, w[name])
        # (quantise in the converted -- chunked -- layout, as the file does, then undo the layout for the oracle)
        d = gq.dequantize(gq.quantize(a, qt), qt, a.shape[1]).astype(np.float32)
        back = {v: k for k, v in enumerate(order)}
        inv = [back[i] for i in range(nv)]

        def unchunk(x, first, blk):
            x = x.copy()
            v = x[first:first + nv * blk].reshape(nv, blk, *x.shape[1:])
            x[first:first + nv * blk] = v[inv].reshape(nv * blk, *x.shape[1:])
            return x
        stem = ".".join(name.split(".")[3:]).rsplit(".", 1)[0] if name.startswith("model.layers.") else ""
        if stem == "linear_attn.in_proj_qkv":
            d = unchunk(d, key2, dv)
        elif stem == "linear_attn.in_proj_z":
            d = unchunk(d, 0, dv)
        elif stem == "linear_attn.out_proj":
            d = np.ascontiguousarray(unchunk(np.ascontiguousarray(d.T), 0, dv).T)
        wq[name] = synth.bf16_round(d)                    # the hybrid's kernels stream bf16: dequantised values are rounded once, at load
    mq = crane_b200.Qwen3_5Model.from_gguf(pq, device=0, max_seq_len=512)
    eq = rel_err(mq.forward_step(ids, 0), Qwen3_5Oracle(cfg, wq).forward(ids, 0).numpy())
    print(f"qwen35 gguf (Q4_K / Q6_K / Q8_0 linears dequantised at load): rel {eq:.1e} against the oracle on the dequantised weights")
    assert eq < 1e-3
    mq.close()


# ---- device-side sampler: top-k total order, penalties, top-p, Gumbel-max ----------------------------------------

def test_topk_kernel_known_answers_and_host_order():
    """The reference's own top-k cases (crane-core/tests/rocm_kernels.rs:96-198) through the kernel: indices are compared bit for bit."""
    from oracle import sampling as smp
    assert crane_b200.op_topk(np.array([0.5, -3.0, 7.25, 1.0, 7.5], np.float32), 5).tolist() == [4, 2, 3, 0, 1]
    ties = (np.arange(240_000) % 4).astype(np.float32) * 0.5          # every candidate equal: only the index orders them
    for k in (1, 40, 64):
        got = crane_b200.op_topk(ties, k)
        assert got.tolist() == [j * 4 + 3 for j in range(k)] and len(set(got.tolist())) == k
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(248_320) * 4).astype(np.float32)         # the Qwen3.5 vocabulary
    for k in (1, 20, 31, 32, 33, 37, 40, 63, 64, 128, 512):
        assert np.array_equal(crane_b200.op_topk(x, k), smp.topk_indices(x, k)), f"k={k}"
    for n in (1, 2, 40, 1023, 1024, 1025, 4095, 4097, 12_289, 65_537):
        y = (rng.standard_normal(n) * 4).astype(np.float32)
        y[rng.integers(0, n, size=max(1, n // 7))] = 1.5                # plenty of exact ties
        for k in (1, 7, 40):
            if k <= n:
                assert np.array_equal(crane_b200.op_topk(y, k), smp.topk_indices(y, k)), f"n={n} k={k}"
    with pytest.raises(crane_b200.CraneB200Error):
        crane_b200.op_topk(x, 513)


def test_sampler_penalties_and_draw_against_oracle():
    from oracle import sampling as smp
    # the reference's literal penalty cases (crane-serve/src/engine/sampling.rs:489-640), read back after the kernel
    cases = [([10.0, -10.0, 3.0], 2.0, 0.0, 0.0, [0, 1]), ([10.0, 10.0], 2.0, 1.0, 0.0, [0, 0, 1]), ([10.0] * 3, 1.0, 0.5, 0.0, [0, 0, 0, 1]),
             ([10.0] * 3, 1.0, 0.0, 0.5, [0, 0, 0, 1]), ([5.0, 4.9], 1.0, 0.1, 0.0, [0, 1, 0, 0, 0, 0]), ([10.0] * 3, 1.0, 0.5, 0.2, [0, 0, 0, 1]),
             ([10.0, 10.0], 1.0, -0.5, 0.0, [0, 0, 1]), ([10.0] * 3, 1.0, 0.0, -0.5, [0, 1]), ([1.0, 2.0, 3.0], 1.1, 0.5, 0.5, [])]
    for lg, rp, fp, pp, ctx in cases:
        tok, after = crane_b200.op_sample(np.array(lg, np.float32), temperature=0.0, repetition_penalty=rp, frequency_penalty=fp,
                                          presence_penalty=pp, context=ctx)
        ref = smp.apply_penalties(lg, rp, fp, pp, ctx)
        assert np.array_equal(after, ref), (lg, after, ref)
        assert tok == int(np.flatnonzero(ref == ref.max())[0])
    # random requests over a real vocabulary: same token as the oracle for every combination of the decision tree
    rng = np.random.default_rng(11)
    V = 151_936
    n_same = 0
    combos = [dict(temperature=0.0), dict(temperature=0.8, top_k=40), dict(temperature=1.0, top_k=64, top_p=0.9), dict(temperature=0.7, top_p=0.8),
              dict(temperature=1.3, top_k=5, top_p=0.5), dict(temperature=1.0, top_k=1), dict(temperature=0.9)]
    for trial in range(28):
        lg = (rng.standard_normal(V) * 3).astype(np.float32)
        kw = dict(combos[trial % len(combos)])
        ctx = rng.integers(0, V, size=64).tolist() + [int(np.argmax(lg))] * 3
        kw.update(repetition_penalty=1.1, frequency_penalty=0.2, presence_penalty=0.1, context=ctx)
        need = V if ("top_k" not in kw and "top_p" not in kw and kw["temperature"] > 0) else 64
        u = rng.uniform(1e-7, 0.999, size=need).astype(np.float32)
        tok, after = crane_b200.op_sample(lg, uniforms=u, **kw)
        ref_tok, ref_after = smp.sample(lg, kw["temperature"], kw.get("top_p"), kw.get("top_k"), 1.1, 0.2, 0.1, ctx, u)
        assert np.array_equal(after, ref_after)
        n_same += int(tok == ref_tok)
        assert tok == ref_tok, (trial, kw, tok, ref_tok)
    print(f"sampler: {n_same}/28 sampled tokens equal to the oracle")
    # without caller uniforms the device draws its own: reproducible per seed, different across seeds
    lg = (rng.standard_normal(V) * 3).astype(np.float32)
    a = [crane_b200.op_sample(lg, temperature=1.0, top_k=50, seed=s)[0] for s in (1, 1, 2, 3, 4, 5)]
    assert a[0] == a[1] and len(set(a[1:])) > 1


def test_sampling_through_the_model_handle():
    """forward_step_sample / sample / topk / decode_batch_sample: logits stay on the device, the ids equal the oracle's on the same logits."""
    from oracle import sampling as smp
    cfg = synth.TINY_QWEN3
    w = dict(synth.synth_checkpoint(cfg))
    m = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=256, max_batch=4)
    m.load_checkpoint(w.items())
    ids = synth.synth_token_ids(20, cfg["vocab_size"], "samp")
    rng = np.random.default_rng(3)
    u = rng.uniform(1e-7, 0.999, size=64).astype(np.float32)
    kw = dict(temperature=0.9, top_k=20, top_p=0.9, repetition_penalty=1.2, context=list(ids[-8:]))
    tok = m.forward_step_sample(ids, 0, uniforms=u, **kw)
    lg = m.copy_logits()                                           # AFTER the penalties (applied in place, as the reference does)
    idx, val = m.topk(20)
    assert np.array_equal(idx, smp.topk_indices(lg, 20)) and np.array_equal(val, lg[idx])
    m.clear_kv_cache()
    raw = m.forward_step(ids, 0)
    assert tok == smp.sample(raw, 0.9, 0.9, 20, 1.2, 0.0, 0.0, list(ids[-8:]), u)[0]
    assert m.sample(temperature=0.0) == int(np.flatnonzero(raw == raw.max())[0])
    # batched: every sequence with its own request
    m.clear_kv_cache()
    seqs, first, raws = [], [], []
    prompts = [synth.synth_token_ids(6 + 5 * i, cfg["vocab_size"], f"sb{i}") for i in range(3)]
    for p in prompts:
        s = m.seq_create()
        m.seq_select(s)
        first.append(m.forward_step_argmax(p, 0))
        seqs.append(s)
    params = [dict(temperature=0.0), dict(temperature=1.0, top_k=8, uniforms=u), dict(temperature=0.7, top_k=30, top_p=0.7, uniforms=u[::-1].copy())]
    got = m.decode_batch_sample(seqs, first, params)
    for i, s in enumerate(seqs):                                   # the same step alone through the single-sequence path
        m2 = crane_b200.Qwen3Model(cfg, device=0, max_seq_len=256)
        m2.load_checkpoint(w.items())
        m2.forward_step(prompts[i], 0)
        raw = m2.forward_step([first[i]], len(prompts[i]))
        pr = params[i]
        assert int(got[i]) == smp.sample(raw, pr["temperature"], pr.get("top_p"), pr.get("top_k"), uniforms=pr.get("uniforms"))[0], f"seq {i}"
        m2.close()
    m.close()


def test_pass_profiler_spans_partition_the_pass():
    """crane_b200_prof_enable / prof_report (ops/prof.rs:37-61): enqueue <= wall, the stage spans carry device time, their sum is
    the pass's device time, and profiling does not change results."""
    cfg = synth.TINY_QWEN3_VL
    m, w = _model(cfg, cls=crane_b200.Qwen3VLModel)
    image = synth.synth_image(96, 64, "prof")
    pv, grid = synth.patchify(image)
    ids = synth.build_vl_prompt(cfg, 30, grid, "prof")
    ref = m.forward(ids, pv, [grid], 0).copy()
    m.clear_kv_cache()
    m.prof_enable(True)
    got = m.forward(ids, pv, [grid], 0)
    toks = m.decode_greedy(int(np.argmax(got)), len(ids), 5)
    rep = m.prof_report()
    m.prof_enable(False)
    assert np.array_equal(got, ref) and len(toks) == 5
    pre, dec = rep["prefill"], rep["decode"]
    assert pre["passes"] == 1 and pre["tokens"] == len(ids) and dec["passes"] == 5
    assert 0 < pre["enqueue_ms"] <= pre["wall_ms"] and pre["device_ms"] > 0
    for name in ("embed", "norm", "attn.qkv", "attn.rope", "attn.flash", "attn.o", "mlp.gate_up", "mlp.down", "head", "vit.patch", "vit.qkv",
                 "vit.flash", "vit.fc1", "vit.merger"):
        assert pre["spans"][name]["device_ms"] > 0, name
    assert abs(sum(v["device_ms"] for v in pre["spans"].values()) - pre["device_ms"]) < 1e-3 * pre["device_ms"] + 1e-3
    assert dec["spans"]["decode"]["device_ms"] > 0
    m.close()
