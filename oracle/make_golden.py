"""Generate tests/golden/*.npz from HF transformers (the reference's stated ground truth,
README.md:401-404) -- run HERE (CPU container), commit the output.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden

For each tiny config in crane_b200.synth the seeded synthetic checkpoint is loaded into the HF
module (float32, eager attention) and into the oracle; the script asserts they agree and stores
HF's outputs as the fixture the oracle (CPU test) and the CUDA path (GPU test) are held to.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from crane_b200 import synth
from oracle.qwen3 import Qwen3Oracle
from oracle.qwen3_vl import Qwen3VLOracle, build_position_ids

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load_hf(model, weights, rename=lambda k: k):
    sd = {rename(k): torch.from_numpy(v) for k, v in weights.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "lm_head" not in m and "inv_freq" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    model.float().eval()
    return model


def golden_qwen3(name, cfg, n_prompt=24, n_new=8):
    from transformers import Qwen3Config, Qwen3ForCausalLM
    w = dict(synth.synth_checkpoint(cfg))
    hf_cfg = Qwen3Config(**{k: v for k, v in cfg.items() if k != "model_type"}, attn_implementation="eager")
    hf = _load_hf(Qwen3ForCausalLM(hf_cfg), w)
    ids = synth.synth_token_ids(n_prompt, cfg["vocab_size"], tag=name)
    orc = Qwen3Oracle(cfg, w)
    with torch.no_grad():
        toks = [int(t) for t in ids]
        hf_logits, tokens = [], []
        past = None
        for step in range(n_new):
            ctx = toks if step == 0 else toks[-1:]
            out = hf(input_ids=torch.tensor([ctx]), past_key_values=past, use_cache=True)
            past = out.past_key_values
            lg = out.logits[0, -1].float()
            lo = orc.forward(ctx, len(toks) - len(ctx))
            err = float((lg - lo).abs().max() / lg.abs().max())
            assert err < 2e-5, f"{name}: oracle vs HF step {step}: {err}"
            nxt = int(lg.argmax())
            assert nxt == int(lo.argmax())
            hf_logits.append(lg.numpy())
            tokens.append(nxt)
            toks.append(nxt)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), prompt=ids, logits=np.stack(hf_logits),
                        tokens=np.array(tokens, np.uint32))
    print(f"{name}: ok, tokens {tokens}")


def golden_qwen3_vl(name, cfg, img_hw=(64, 96), n_text=20, n_new=6):
    from transformers.models.qwen3_vl import Qwen3VLConfig, Qwen3VLForConditionalGeneration
    w = dict(synth.synth_checkpoint(cfg))
    tc = dict(cfg["text_config"])
    rs = tc.pop("rope_scaling")
    tc["rope_parameters"] = {"rope_type": "default", "rope_theta": tc.pop("rope_theta"), **rs}
    hf_cfg = Qwen3VLConfig(text_config=tc, vision_config=dict(cfg["vision_config"]),
                           image_token_id=cfg["image_token_id"],
                           vision_start_token_id=cfg["vision_start_token_id"],
                           vision_end_token_id=cfg["vision_end_token_id"],
                           tie_word_embeddings=True)
    hf_cfg._attn_implementation = "eager"
    hf_cfg.text_config._attn_implementation = "eager"
    hf_cfg.vision_config._attn_implementation = "eager"
    hf = _load_hf(Qwen3VLForConditionalGeneration(hf_cfg), w)
    image = synth.synth_image(*img_hw, tag=name)
    pv, grid = synth.patchify(image)
    ids = synth.build_vl_prompt(cfg, n_text, grid, tag=name)
    # HF's activation choice (tanh in the ViT MLP, erf in the mergers) for the cross-check
    orc = Qwen3VLOracle(cfg, w, vit_act="tanh", merger_act="erf")
    ref = Qwen3VLOracle(cfg, w)       # the reference's choice: what the fixture's `ref_*` hold
    with torch.no_grad():
        vis = hf.model.visual(torch.from_numpy(pv), grid_thw=torch.tensor([grid]))
        hf_img = vis.pooler_output if hasattr(vis, "pooler_output") else vis[0]
        o_img, o_deep = orc.vision.forward(pv, [grid])
        err = float((hf_img - o_img).abs().max() / hf_img.abs().max())
        assert err < 2e-5, f"{name}: ViT oracle vs HF: {err}"
        r_img, r_deep = ref.vision.forward(pv, [grid])
        pos3, nxt = build_position_ids(ids, [grid], 2, cfg["image_token_id"])
        tids = torch.tensor([ids.astype(np.int64)])
        mm = (tids == cfg["image_token_id"]).int()
        hf_pos, _ = hf.model.get_rope_index(tids, mm, image_grid_thw=torch.tensor([grid]))
        assert np.array_equal(hf_pos[:, 0].numpy(), pos3.astype(np.int64)), "position ids differ from HF"
        out = hf(input_ids=tids, pixel_values=torch.from_numpy(pv), mm_token_type_ids=mm,
                 image_grid_thw=torch.tensor([grid]), use_cache=True)
        past = out.past_key_values
        lg = out.logits[0, -1].float()
        lo = orc.prefill(ids, pv, [grid])
        err = float((lg - lo).abs().max() / lg.abs().max())
        assert err < 2e-5, f"{name}: VL prefill oracle vs HF: {err}"
        hf_logits, tokens = [lg.numpy()], [int(lg.argmax())]
        ref_logits = [ref.prefill(ids, pv, [grid]).numpy()]
        ref_tokens = [int(ref_logits[0].argmax())]
        S = len(ids)
        for step in range(1, n_new):
            p = nxt + step - 1
            out = hf(input_ids=torch.tensor([[tokens[-1]]]), past_key_values=past, use_cache=True,
                     position_ids=torch.full((3, 1, 1), p))
            past = out.past_key_values
            lg = out.logits[0, -1].float()
            lo = orc.decode_step(tokens[-1], S + step - 1)
            err = float((lg - lo).abs().max() / lg.abs().max())
            assert err < 2e-5, f"{name}: VL decode {step} oracle vs HF: {err}"
            hf_logits.append(lg.numpy())
            tokens.append(int(lg.argmax()))
            rl = ref.decode_step(ref_tokens[-1], S + step - 1).numpy()
            ref_logits.append(rl)
            ref_tokens.append(int(rl.argmax()))
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"), image=image, prompt=ids, grid=np.array(grid), pos3=pos3,
        hf_image_embeds=hf_img.numpy(), hf_logits=np.stack(hf_logits), hf_tokens=np.array(tokens, np.uint32),
        ref_image_embeds=r_img.numpy(), ref_deepstack=np.stack([d.numpy() for d in r_deep]),
        ref_logits=np.stack(ref_logits), ref_tokens=np.array(ref_tokens, np.uint32))
    print(f"{name}: ok, hf tokens {tokens}, ref-act tokens {ref_tokens}")


def golden_qwen3_5(name, cfg, n_prompt=21, n_new=6):
    from transformers.models.qwen3_5 import Qwen3_5ForCausalLM, Qwen3_5TextConfig
    from oracle.qwen3_5 import Qwen3_5Oracle
    w = dict(synth.synth_checkpoint(cfg))
    hf_cfg = Qwen3_5TextConfig(**{k: v for k, v in cfg.items() if k != "model_type"})
    hf_cfg._attn_implementation = "eager"
    hf = _load_hf(Qwen3_5ForCausalLM(hf_cfg), w)
    ids = synth.synth_token_ids(n_prompt, cfg["vocab_size"], tag=name)
    orc = Qwen3_5Oracle(cfg, w)
    with torch.no_grad():
        toks = [int(t) for t in ids]
        hf_logits, tokens = [], []
        past = None
        for step in range(n_new):
            ctx = toks if step == 0 else toks[-1:]
            out = hf(input_ids=torch.tensor([ctx]), past_key_values=past, use_cache=True)
            past = out.past_key_values
            lg = out.logits[0, -1].float()
            lo = orc.forward(ctx, len(toks) - len(ctx))
            err = float((lg - lo).abs().max() / lg.abs().max())
            assert err < 5e-5, f"{name}: oracle vs HF step {step}: {err}"
            nxt = int(lg.argmax())
            hf_logits.append(lg.numpy())
            tokens.append(nxt)
            toks.append(nxt)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), prompt=ids, logits=np.stack(hf_logits),
                        tokens=np.array(tokens, np.uint32))
    print(f"{name}: ok, tokens {tokens}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    golden_qwen3("tiny_qwen3", synth.TINY_QWEN3)
    golden_qwen3("tiny_qwen3_untied", synth.TINY_QWEN3_UNTIED)
    golden_qwen3_vl("tiny_qwen3_vl", synth.TINY_QWEN3_VL)
    golden_qwen3_5("tiny_qwen3_5", synth.TINY_QWEN3_5)


if __name__ == "__main__":
    main()
