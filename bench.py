#!/usr/bin/env python
"""bench.py -- Qwen3-VL-2B decode tok/s + prefill TFLOP/s on N x B200 (BASELINE.json `metric`), with the
kernel roofline and the reference-equivalent CPU path timed beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU arithmetic (oracle) on host cores

A *step* is one request of BASELINE.json configs[1]: one 448x448 image + 256 prompt tokens
(ViT -> splice -> prefill of 454 positions) followed by 256 greedy decode tokens.  Weights are the seeded
synthetic checkpoint of crane_b200/synth.py (no checkpoints offline); per-GPU work is fixed as N grows
(independent requests per rank, no data-path collective): "scaling": "weak".

Only the `cpu_baseline` leg and `--impl reference` import oracle/ (as the thing being timed, never as a
fallback for the CUDA path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from crane_b200 import synth  # noqa: E402

N_TEXT = 256
N_DECODE = 256
IMAGE_HW = (448, 448)


# --------------------------------------------------------------------------------------------------
# workload + algorithmic cost model
# --------------------------------------------------------------------------------------------------
def make_request(cfg, tag="bench"):
    image = synth.synth_image(*IMAGE_HW, tag=tag)
    pv, grid = synth.patchify(image)
    ids = synth.build_vl_prompt(cfg, N_TEXT, grid, tag=tag)
    return ids, pv, grid


def decode_bytes_per_token(cfg, ctx):
    tc = cfg["text_config"]
    H, I, L, V = tc["hidden_size"], tc["intermediate_size"], tc["num_hidden_layers"], tc["vocab_size"]
    nh, nkv, d = tc["num_attention_heads"], tc["num_key_value_heads"], tc["head_dim"]
    per_layer = ((nh + 2 * nkv) * d * H + H * nh * d + 2 * I * H + H * I) * 2
    weights = per_layer * L + V * H * 2
    kv = (ctx * 2 * nkv * d * 2 + 2 * nkv * d * 2) * L            # read ctx cached tokens, append one
    return weights + kv


def prefill_flops(cfg, S, n_patches, n_img_tok):
    tc, vc = cfg["text_config"], cfg["vision_config"]
    H, I, L, V = tc["hidden_size"], tc["intermediate_size"], tc["num_hidden_layers"], tc["vocab_size"]
    nh, nkv, d = tc["num_attention_heads"], tc["num_key_value_heads"], tc["head_dim"]
    lin = ((nh + 2 * nkv) * d * H + H * nh * d + 2 * I * H + H * I) * L
    text = 2 * lin * S + 4 * S * (S / 2) * nh * d * L + 2 * V * H
    Hv, Iv, Lv = vc["hidden_size"], vc["intermediate_size"], vc["depth"]
    pk = vc["in_channels"] * vc["temporal_patch_size"] * vc["patch_size"] ** 2
    mh = Hv * vc["spatial_merge_size"] ** 2
    vit_lin = (3 * Hv * Hv + Hv * Hv + 2 * Hv * Iv) * Lv
    n_merg = 1 + len(vc.get("deepstack_visual_indexes", []))
    vit = 2 * vit_lin * n_patches + 4 * n_patches * n_patches * Hv * Lv + 2 * pk * Hv * n_patches \
        + n_merg * 2 * (mh * mh + mh * vc["out_hidden_size"]) * n_img_tok
    return text + vit


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def synth_checkpoint_parallel(cfg, as_bits):
    """Seeded synthetic checkpoint, tensors drawn in parallel; bf16 bit patterns for the GPU arm, f32 for the oracle."""
    import concurrent.futures as cf
    specs = list(synth.tensor_specs(cfg))

    def one(spec):
        name, shape, kind = spec
        x = synth.make_tensor(name, shape, kind)
        return name, (synth.f32_to_bf16_bits(x) if as_bits and x.ndim >= 2 and x.size > 1 << 16 else x)
    with cf.ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        return list(ex.map(one, specs))


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's f32 path on the host cores
# --------------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Threads the CPU arm uses: the cores this process may actually run on (affinity mask and cgroup quota), capped at 32
    -- past that torch's f32 GEMV/GEMM on this model stops scaling and, on an over-subscribed 128-vCPU box, collapses."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, 32))


def run_cpu(cfg, weights, n_decode, steps, warmup, budget_s=None, keep_logits=False):
    """`warmup` + `steps` requests (ViT + prefill + n_decode greedy tokens) on the host cores.  With `budget_s` the number of
    timed steps is cut so the whole call ends inside it (projected from the first request) -- what actually ran is returned."""
    import torch
    from oracle.qwen3_vl import Qwen3VLOracle
    cores = host_threads()
    torch.set_num_threads(cores)
    orc = Qwen3VLOracle(cfg, {k: v for k, v in weights}, max_pos=2048)
    ids, pv, grid = make_request(cfg)
    pre_t, dec_t, toks = [], [], 0
    first_logits, tok_list, step_logits = None, [], []
    t_start = time.perf_counter()
    it, warm_run, steps_run = 0, 0, 0
    with torch.no_grad():
        while steps_run < steps:
            orc.clear_kv_cache()
            t0 = time.perf_counter()
            lg = orc.prefill(ids, pv, [grid])
            t1 = time.perf_counter()
            tok = int(lg.argmax())
            first_logits, tok_list, step_logits = lg.numpy().copy(), [tok], []
            for i in range(n_decode):
                lg = orc.decode_step(tok, len(ids) + i)
                if keep_logits:
                    step_logits.append(lg.numpy().copy())
                tok = int(lg.argmax())
                tok_list.append(tok)
            t2 = time.perf_counter()
            if it >= warmup:
                pre_t.append(t1 - t0)
                dec_t.append(t2 - t1)
                toks += n_decode
                steps_run += 1
            else:
                warm_run += 1
            it += 1
            if budget_s is not None:
                per = t2 - t0
                left = budget_s - (time.perf_counter() - t_start)
                if it < warmup and left < per * (warmup - it + 1):
                    warmup = it                                     # no room for more warm-up: start timing now
                if steps_run >= 1 and left < per:
                    break
    n_patches = pv.shape[0]
    fl = prefill_flops(cfg, len(ids), n_patches, n_patches // 4)
    return {"decode_tok_s": toks / sum(dec_t), "request_tok_s": toks / (sum(dec_t) + sum(pre_t)),
            "prefill_tflops": fl / (sum(pre_t) / len(pre_t)) / 1e12,
            "prefill_s": sum(pre_t) / len(pre_t), "cores": cores, "ms_per_step": 1e3 * (sum(pre_t) + sum(dec_t)) / steps_run,
            "steps_run": steps_run, "warmup_run": warm_run, "n_decode": n_decode,
            "prefill_logits": first_logits, "tokens": tok_list, "step_logits": step_logits}


def bench_config4(args, rank, world, local_rank):
    """BASELINE.json configs[3]: Qwen3-8B with a Q4_K_M-like GGUF recipe, 32 sequences decoded in lock-step, sharded over the N
    GPUs of the box (32 / N per rank: the total batch is fixed, "scaling": "strong"); every round ends with the NCCL all-gather of
    the 32 logits rows and greedy tokens (crane_b200_decode_batch_gather).  A step = one decode round of the whole batch.
    Weights are syntactically valid random ggml blocks (tools/bench_configs.py fake_blocks: timing only, parity of the quantised
    path is tests/test_gpu_parity.py); the CPU arm of this config is not part of the driver's contract and is not run."""
    import torch
    import crane_b200
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as bc
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "config 4's CPU arm is not timed (an 8B f32 oracle decode of 32 sequences does not fit the bench budget); the headline config has the reference arm"}))
        return
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    n_total, prompt_len = 32, 128
    mine = crane_b200.shard_sequences(n_total, world, rank)
    cfg = synth.QWEN3_8B
    quant = {"q_proj.weight": "Q4_K", "k_proj.weight": "Q4_K", "v_proj.weight": "Q6_K", "o_proj.weight": "Q4_K", "gate_proj.weight": "Q4_K",
             "up_proj.weight": "Q4_K", "down_proj.weight": "Q6_K", "lm_head.weight": "Q6_K", "embed_tokens.weight": "Q4_K"}
    m = crane_b200.Qwen3Model(cfg, device=dev, max_seq_len=1024, max_batch=len(mine))
    bc.load_cheap(m, cfg, quant)
    uid = [crane_b200.comm_unique_id() if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(uid, src=0)
    m.comm_init(uid[0], rank, world)
    slots, toks = [], []
    for j, g in enumerate(mine):
        s = 0 if j == 0 else m.seq_create()
        m.seq_select(s)
        toks.append(m.forward_step_argmax(synth.synth_token_ids(prompt_len, cfg["vocab_size"], f"c4-{g}"), 0))
        slots.append(s)
    sampler = ClockSampler(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    import ctypes
    C_byref = ctypes.byref
    gather_logits = os.environ.get("CRANE_B200_C4_GATHER", "logits") == "logits"
    lg = crane_b200.Logits()
    seqs_np = np.ascontiguousarray(slots, dtype=np.int32)
    out_np = np.empty(world * len(mine), dtype=np.uint32)

    def one_round_raw(toks):          # the C call itself: logits rows stay on the device (all-gathered there), 4 bytes per sequence return
        t = np.ascontiguousarray(toks, dtype=np.uint32)
        m._ck(m.lib.crane_b200_decode_batch_gather(m.h, crane_b200._ptr(seqs_np), crane_b200._ptr(t), seqs_np.size, crane_b200._ptr(out_np),
                                                   C_byref(lg) if gather_logits else None))
        return [int(x) for x in out_np[rank * len(mine):(rank + 1) * len(mine)]]

    for _ in range(max(args.warmup, 3)):
        toks = one_round_raw(toks)
    barrier()
    sampler.start()
    l0 = m.kernel_launches()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        toks = one_round_raw(toks)
        dev_ms += m.last_timing()["decode_ms"]
    barrier()
    wall = time.perf_counter() - t0
    launches = m.kernel_launches() - l0
    clocks = sampler.summary()
    stats = torch.tensor([wall, dev_ms / 1e3], dtype=torch.float64, device=f"cuda:{dev}")
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    wall_m, dev_m = [float(x) for x in stats.tolist()]
    if rank == 0:
        H, I, L, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"], cfg["vocab_size"]
        q4 = (32 * 128 * H + 8 * 128 * H + H * 32 * 128 + 2 * I * H) * 0.5625
        q6 = (8 * 128 * H + H * I) * 0.875
        groups = (len(mine) + 3) // 4                      # weight passes per round on one rank (<= 4 sequences share one)
        ctx = prompt_len + max(args.warmup, 3) + args.steps / 2
        by = groups * ((q4 + q6) * L + V * H * 0.875) + len(mine) * ctx * 2 * 8 * 128 * 2 * 2 * L
        peaks = load_peaks()
        step_s = dev_m / args.steps
        out = {"metric": "decode_tok_per_s", "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "q4_k/q6_k weights x q8_k activations (int dot), f32 accumulate",
               "data": "synthetic",
               "config": {"workload": f"Qwen3-8B Q4_K_M-like GGUF recipe, batch {n_total} lock-step greedy decode after {prompt_len}-token prompts, "
                                      f"{len(mine)} sequences per GPU", "weights": "random ggml blocks (timing only)",
                          "l2": "per-round weight stream 4.9 GB >> 126 MB L2", "parallelism": f"dp{world}: sequences sharded, NCCL all-gather of "
                                      + ("logits [32, V] + tokens" if gather_logits else "tokens") + " per round"},
               "value": n_total * args.steps / wall_m, "ms_per_step": 1e3 * wall_m / args.steps,
               "e2e": {"value": n_total * args.steps / wall_m, "unit": "tok/s", "h2d_bytes_per_step": int(len(mine) * (4 + 32)),
                       "d2h_bytes_per_step": int(n_total * 4), "api": "crane_b200_decode_batch_gather (host token ids in, gathered token ids out)"},
               "device_ms_per_step": 1e3 * step_s, "gpu_launches": int(launches), "clocks": clocks,
               "roofline": {"bound": "hbm", "achieved": by / step_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / step_s / 1e9 / peaks["hbm_gbs"],
                            "traffic": None, "peak_source": peaks["source"], "kernel": "cb::qgemv_kernel (Q4_K / Q6_K x Q8_K integer-dot GEMV, 4 sequences per weight pass)",
                            "algorithmic_bytes_per_launch": by, "launch": f"one decode round on one rank: {groups} weight pass(es) + KV of {len(mine)} sequences"}}
        print(json.dumps(out))
    m.close()
    if dist is not None:
        dist.destroy_process_group()


def bench_config2(args, rank, world, local_rank):
    """BASELINE.json configs[2]: Qwen3.5-0.8B (3 Gated-Delta-Net layers per gated full-attention layer) bf16, one request =
    4096-token prefill + 512 greedy decode tokens on one GPU; N > 1 = N independent replicas ("scaling": "weak").  A step = one request.
    `value` = decode tok/s (CUDA events around the on-device loop), `prefill_*` from the events around the prefill pass, `e2e` =
    decode tokens / wall time of the whole request through the generate call (host token ids in, host token ids out).
    Weights are cheap tiled random bf16 blocks (timing only: parity of every kernel on this path is tests/test_gpu_parity.py, incl. the
    full-width geometry and chunkwise-vs-sequential recurrence tests); the CPU arm of this config is not timed."""
    import torch
    import crane_b200
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as bc
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "config 2's CPU arm is not timed (a 4096-token f32 oracle prefill with a python-loop recurrence does not fit the bench budget); the headline config has the reference arm"}))
        return
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    S, n_dec = 4096, 512
    cfg = synth.QWEN3_5_0_8B
    m = crane_b200.Qwen3_5Model(cfg, device=dev, max_seq_len=S + n_dec + 128)
    bc.load_cheap(m, cfg)
    ids = synth.synth_token_ids(S, cfg["vocab_size"], f"c2-{rank}")
    sampler = ClockSampler(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def request():
        m.clear_kv_cache()
        t0 = time.perf_counter()
        out = m.generate_greedy(ids, n_dec + 1)          # first token from the prefill's argmax, then n_dec decode steps on the device
        wall = time.perf_counter() - t0
        t = m.last_timing()
        return wall, t["prefill_ms"], t["decode_ms"], len(out) - 1

    for _ in range(max(args.warmup, 3)):
        request()
    barrier()
    sampler.start()
    l0 = m.kernel_launches()
    walls, pres, decs, ntok = [], [], [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        w_, p_, d_, n_ = request()
        walls.append(w_); pres.append(p_); decs.append(d_); ntok += n_
    barrier()
    wall_all = time.perf_counter() - t0
    launches = m.kernel_launches() - l0
    clocks = sampler.summary()
    stats = torch.tensor([wall_all, sum(decs) / 1e3, sum(pres) / 1e3, sum(walls)], dtype=torch.float64, device=f"cuda:{dev}")
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    wall_m, dec_m, pre_m, req_m = [float(x) for x in stats.tolist()]
    if rank == 0:
        H, I, L, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"], cfg["vocab_size"]
        full, gdn = L // 4, L - L // 4
        w_full = (8 * 256 * 2 + 2 * 2 * 256) * H + H * 8 * 256
        w_gdn = (6144 + 2048 + 32) * H + H * 2048
        ctx = S + n_dec / 2
        by = ((w_full * full + w_gdn * gdn + 3 * I * H * L) + V * H) * 2 + gdn * 2 * 16 * 128 * 128 * 4 + ctx * 2 * 2 * 256 * 2 * 2 * full
        fl = 2 * (w_full * full + w_gdn * gdn + 3 * I * H * L) * S + 4 * S * (S / 2) * 8 * 256 * full + 2 * V * H
        peaks = load_peaks()
        n_steps_total = world * ntok
        step_s = dec_m / ntok                               # seconds per decode step on the slowest rank
        out = {"metric": "decode_tok_per_s", "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": f"Qwen3.5-0.8B hybrid (18 Gated-Delta-Net + 6 gated full-attention layers) bf16, {S}-token prefill + {n_dec} greedy decode tokens per request",
                          "weights": "random bf16 blocks (timing only)", "l2": "per-token weight stream 1.5 GB >> 126 MB L2",
                          "parallelism": f"dp{world} (replicas)", "gdn": os.environ.get("CRANE_B200_GDN", "auto (chunkwise recurrence for the prefill)")},
               "value": n_steps_total / dec_m, "ms_per_step": 1e3 * wall_m / args.steps,
               "prefill_ms": 1e3 * pre_m / args.steps, "prefill_tflops": world * fl * args.steps / pre_m / 1e12,
               "prefill_frac_of_bf16_peak": fl * args.steps / pre_m / 1e12 / peaks["bf16_tflops"], "prefill_tok_per_s": world * S * args.steps / pre_m,
               "e2e": {"value": n_steps_total / req_m, "unit": "tok/s", "h2d_bytes_per_step": int(S * 4), "d2h_bytes_per_step": int((n_dec + 1) * 4),
                       "api": "crane_b200_generate_greedy (host prompt ids in, host token ids out; prefill inside the timed region)"},
               "gpu_launches": int(launches), "clocks": clocks,
               "roofline": {"bound": "hbm", "achieved": by / step_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / step_s / 1e9 / peaks["hbm_gbs"],
                            "traffic": None, "peak_source": peaks["source"], "kernel": "decode step of the hybrid model (gemv_kernel chain + gdn_decode_kernel + attn_decode_kernel<256,4>)",
                            "algorithmic_bytes_per_launch": by, "launch": "one decode step: every weight once + 18 recurrent states read and written + KV of 6 layers"}}
        print(json.dumps(out))
    m.close()
    if dist is not None:
        dist.destroy_process_group()


def other_config_summary(config, steps=2, timeout_s=120):
    """One more BASELINE config measured by THIS script in a child process (`bench.py --config <config>`), condensed for the headline
    line's `other_configs` field.  A child, so that nothing it does -- a failure, a hang (killed at the timeout) -- can keep the
    headline line from being printed; whatever goes wrong comes back as {"error": ...}."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", config, "--steps", str(steps), "--warmup", "3", "--gpus", "1"],
                           capture_output=True, text=True, timeout=timeout_s, env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not line:
            return {"error": f"rc {r.returncode}: {(r.stderr or r.stdout).strip()[-200:]}"}
        d = json.loads(line[-1])
        keep = ("value", "unit", "steps", "ms_per_step", "prefill_ms", "prefill_tflops", "prefill_frac_of_bf16_peak", "prefill_tok_per_s", "gpu_launches", "clocks")
        out = {k: d[k] for k in keep if k in d}
        out["workload"] = d.get("config", {}).get("workload")
        out["e2e_tok_s"] = d.get("e2e", {}).get("value")
        out["decode_roofline_frac"] = d.get("roofline", {}).get("frac")
        out["how"] = f"child process: bench.py --config {config} --steps {steps} --warmup 3 (same GPU, after the headline measurement)"
        return out
    except Exception as e:                                 # incl. subprocess.TimeoutExpired
        return {"error": repr(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="crane_b200", choices=["crane_b200", "reference"])
    ap.add_argument("--config", default="qwen3_vl_2b", choices=["qwen3_vl_2b", "tiny", "qwen3_8b_q4km", "qwen3_5_0_8b"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the config-2 (Qwen3.5-0.8B) child measurement appended to the headline line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == "qwen3_8b_q4km":
        return bench_config4(args, rank, world, local_rank)
    if args.config == "qwen3_5_0_8b":
        return bench_config2(args, rank, world, local_rank)
    cfg = synth.QWEN3_VL_2B if args.config == "qwen3_vl_2b" else synth.TINY_QWEN3_VL
    name = "Qwen3-VL-2B" if args.config == "qwen3_vl_2b" else "tiny-Qwen3-VL"
    workload = f"{name} bf16, 1x(448x448) image + {N_TEXT}-tok prompt, prefill + {N_DECODE} greedy decode tokens per request"
    base = {"metric": "decode_tok_per_s", "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "weights": "seeded synthetic N(0,1/fan_in) rounded to bf16",
                       "l2": "per-token weight stream 3.44 GB >> 126 MB L2 (inputs larger than L2)", "parallelism": f"dp{args.gpus} (replicas)",
                       "precision": os.environ.get("CRANE_B200_PRECISION", "split") + " (split = hi+lo bf16 operands / KV pages, parity mode)"}}

    # ---------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        w = synth_checkpoint_parallel(cfg, as_bits=False)
        # Every step is the whole request (ViT + prefill + 256 greedy tokens), as on the CUDA arm; --steps / --warmup are honoured up
        # to a wall-clock budget (a request takes ~20 s on 16 threads), and the line carries what actually ran.
        budget = float(os.environ.get("CRANE_B200_REF_BUDGET_S", "170"))
        r = run_cpu(cfg, w, N_DECODE, max(1, args.steps), max(0, args.warmup), budget_s=budget)
        sample = (f"{r['warmup_run']} warm-up + {r['steps_run']} timed full requests (ViT + prefill + {N_DECODE} decode tokens), torch f32 on "
                  f"{r['cores']} threads (oracle port of the Candle CPU path); asked for --steps {args.steps} --warmup {args.warmup}, cut by a {budget:.0f} s budget")
        out = dict(base, impl="reference", value=r["decode_tok_s"], ms_per_step=r["ms_per_step"], prefill_tflops=r["prefill_tflops"],
                   steps=r["steps_run"], warmup=r["warmup_run"], steps_requested=args.steps, warmup_requested=args.warmup,
                   dtype="f32", cpu_baseline={"value": r["decode_tok_s"], "unit": "tok/s", "cores": r["cores"], "kind": "port", "sample": sample},
                   e2e={"value": r["request_tok_s"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "definition": "decode tokens / (ViT + prefill + decode) wall time of the request", "decode_only_tok_s": r["decode_tok_s"],
                        "prefill_s": r["prefill_s"]},
                   gpu_launches=0, n_gpus=args.gpus)
        print(json.dumps(out))
        return

    # ---------------------------------------------------------------------------- CUDA arm
    import torch
    import crane_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the crane_b200 path has no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    sampler = ClockSampler(dev)

    weights = synth_checkpoint_parallel(cfg, as_bits=True)
    model = crane_b200.Qwen3VLModel(cfg, device=dev, max_seq_len=1024)
    model.load_checkpoint(weights)
    ids, pv, grid = make_request(cfg)
    S, n_patches = len(ids), pv.shape[0]
    n_img_tok = n_patches // 4
    pin = torch.from_numpy(pv).pin_memory()          # pinned host staging for the e2e leg
    pv_pinned = pin.numpy()

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident leg: prefill via the public call, decode with the on-device greedy loop ----
    gpu_out = {}

    def request_device():
        model.clear_kv_cache()
        lg = model.forward(ids, pv_pinned, [grid], 0)
        first = int(np.argmax(lg))
        toks = model.decode_greedy(first, S, N_DECODE)
        t = model.last_timing()
        gpu_out["prefill_logits"], gpu_out["tokens"] = lg, [first] + [int(x) for x in toks]
        return t["prefill_ms"], t["decode_ms"], toks

    for _ in range(args.warmup):
        request_device()
    barrier()
    sampler.start()
    l0 = model.kernel_launches()
    pre_ms, dec_ms = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p, d, toks = request_device()
        pre_ms.append(p)
        dec_ms.append(d)
    barrier()
    wall = time.perf_counter() - t0
    launches = model.kernel_launches() - l0

    # ---- e2e leg: the request through the public API with HOST buffers: image + prompt ids in (H2D inside), 256 token ids out.
    # (a) the call a user makes to generate: vl_forward + ONE decode_greedy call (`Model::generate`, qwen3/model.rs:275-349), tokens
    #     read back at the end; (b) the server's streaming path: one vl_decode_step_argmax call (4-byte D2H + host sync) per token.
    def request_e2e(per_token):
        model.clear_kv_cache()
        t_a = time.perf_counter()
        lg = model.forward(ids, pv_pinned, [grid], 0)
        tok = int(np.argmax(lg))
        t_b = time.perf_counter()
        if per_token:
            toks_e = [tok]
            for i in range(N_DECODE):
                tok = model.decode_step_argmax(tok, S + i)
                toks_e.append(tok)
        else:
            toks_e = [tok] + [int(x) for x in model.decode_greedy(tok, S, N_DECODE)]
        gpu_out["tokens_e2e_stream" if per_token else "tokens_e2e"] = toks_e
        return t_b - t_a, time.perf_counter() - t_b

    request_e2e(False)
    request_e2e(True)
    barrier()
    e_pre, e_dec, s_dec = [], [], []
    for _ in range(max(1, min(args.steps, 3))):
        a, b = request_e2e(False)
        e_pre.append(a)
        e_dec.append(b)
        s_dec.append(request_e2e(True)[1])
    barrier()
    clocks = sampler.summary()

    dec_total_s = sum(dec_ms) / 1e3
    stats = torch.tensor([wall, dec_total_s, sum(pre_ms) / 1e3, sum(e_dec), sum(e_pre), sum(s_dec)], dtype=torch.float64, device=f"cuda:{dev}")
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    wall_m, dec_m, pre_m, edec_m, epre_m, sdec_m = [float(x) for x in stats.tolist()]
    tokens_all = world * args.steps * N_DECODE
    value = tokens_all / dec_m
    fl = prefill_flops(cfg, S, n_patches, n_img_tok)
    prefill_tflops = world * args.steps * fl / pre_m / 1e12
    e2e_value = world * len(e_dec) * N_DECODE / (edec_m + epre_m)    # the request end to end: image + ids in, 256 tokens out
    e2e_decode_only = world * len(e_dec) * N_DECODE / edec_m
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    traffic = None           # DRAM bytes of one decode launch measured by ncu (tools/profile.sh + tools/summarize_ncu.py traffic)
    for tf in ("r02_traffic.json", "r01_traffic.json"):
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", tf)))
            break
        except Exception:
            pass
    persistent = model.decode_path() == "persistent"
    # one launch of the dominant kernel: the persistent decode kernel runs all N_DECODE steps (contexts S .. S + N_DECODE - 1) in one
    # launch; the kernel-chain path replays one graph per step
    bytes_launch = sum(decode_bytes_per_token(cfg, S + i) for i in range(N_DECODE))
    bytes_tok = bytes_launch / N_DECODE
    launch_s = (sum(dec_ms) / 1e3) / args.steps                   # this rank's device time per N_DECODE-step decode (CUDA events, engine stream)
    achieved = bytes_launch / launch_s / 1e9
    out = dict(base)
    out.update({
        "value": value, "ms_per_step": 1e3 * wall_m / args.steps,
        "prefill_tflops": prefill_tflops, "prefill_ms": 1e3 * pre_m / args.steps, "prefill_tflop_per_request": fl / 1e12,
        "prefill_frac_of_bf16_peak": prefill_tflops / world / (peaks["bf16_tflops"]),
        "e2e": {"value": e2e_value, "unit": "tok/s", "h2d_bytes_per_step": int(pv.nbytes + ids.nbytes + 3 * 4 * S + 2 * 32),
                "d2h_bytes_per_step": int(4 * N_DECODE + 4 * cfg["text_config"]["vocab_size"]),
                "definition": "decode tokens / (ViT + prefill + decode) wall time of the request, host buffers in and out",
                "decode_only_tok_s": e2e_decode_only, "prefill_s": epre_m / len(e_pre),
                "api": "crane_b200_vl_forward + one crane_b200_decode_greedy call (Model::generate); tokens read back at the end",
                "streaming_tok_s": world * len(s_dec) * N_DECODE / (sdec_m + epre_m),
                "streaming_api": "crane_b200_vl_forward + crane_b200_vl_decode_step_argmax per token (host sync + 4-byte D2H each)",
                "tokens_equal_device_loop": gpu_out["tokens_e2e"] == gpu_out["tokens"] and gpu_out["tokens_e2e_stream"] == gpu_out["tokens"]},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": (traffic.get("decode_launch_dram_bytes") or traffic.get("decode_step_dram_bytes")) if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None, "peak_source": peaks["source"],
                     "kernel": ("cb::decode_ll_kernel: ONE persistent cooperative launch = %d decode steps (all layers, attention, lm_head, argmax, "
                                "next embedding); activations cross CTAs as tagged (value, tag) pairs, no grid barrier" % N_DECODE) if persistent else
                               "decode step = 5 GEMV/attention launches x 28 layers + lm_head (cb::gemv_kernel dominates), one graph replay per step",
                     "algorithmic_bytes_per_launch": bytes_launch if persistent else bytes_tok,
                     "algorithmic_bytes_per_token": bytes_tok,
                     "launch": ("%d decode steps, contexts %d..%d" % (N_DECODE, S, S + N_DECODE - 1)) if persistent else "one decode step (graph replay), mean ctx %d" % (S + N_DECODE // 2),
                     "launch_ms": 1e3 * launch_s if persistent else 1e3 * launch_s / N_DECODE},
    })
    if not args.no_cpu_baseline and world == 1:
        w32 = [(n, synth.bf16_bits_to_f32(a) if a.dtype == np.uint16 else a) for n, a in weights]
        n_par = 64           # bounded sample: one request with 64 decode tokens (~10-15 s), also the full-size parity reference
        r = run_cpu(cfg, w32, n_par, 1, 0, keep_logits=True)
        out["cpu_baseline"] = {"value": r["decode_tok_s"], "unit": "tok/s", "cores": r["cores"], "kind": "port",
                               "prefill_tflops": r["prefill_tflops"],
                               "sample": f"1 timed request, no warm-up: ViT+prefill({S}) + {n_par} decode tokens, torch f32, {r['cores']} threads"}
        # Full-size parity on the same request.  (1) free-running: the GPU's greedy tokens against the oracle's; (2) teacher-forced on
        # the oracle's tokens: logits of EVERY decode step, so one near-tie cannot hide what follows it; (3) the oracle's smallest
        # top-1 / top-2 gap along the way, which says how much a logit may move before a greedy token changes.
        ref, got = r["prefill_logits"].reshape(-1), np.asarray(gpu_out["prefill_logits"]).reshape(-1)
        n_cmp = min(len(r["tokens"]), len(gpu_out["tokens"]))
        same = 0
        while same < n_cmp and gpu_out["tokens"][same] == r["tokens"][same]:
            same += 1
        model.clear_kv_cache()
        model.forward(ids, pv_pinned, [grid], 0)
        rels, margins = [], []
        for i in range(n_par):
            lg = np.asarray(model.decode_step(r["tokens"][i], S + i)).reshape(-1)
            o = r["step_logits"][i].reshape(-1)
            rels.append(float(np.abs(lg - o).max() / np.abs(o).max()))
            top2 = np.partition(o, -2)[-2:]
            margins.append(float((top2[1] - top2[0]) / np.abs(o).max()))
        out["parity"] = {"prefill_logits_rel": float(np.abs(got - ref).max() / np.abs(ref).max()),
                         "greedy_tokens_equal_prefix": f"{same}/{n_cmp}",
                         "teacher_forced_steps": n_par, "step_logits_rel_max": max(rels), "last_step_logits_rel": rels[-1],
                         "min_top2_margin_rel": min(margins), "min_top2_margin_at_step": int(np.argmin(margins)),
                         "against": "oracle (torch f32) on the same full-size request; rel = max|d| / max|ref|"}
    if world == 1 and args.config == "qwen3_vl_2b" and not args.no_cpu_baseline and not args.no_other_configs and os.environ.get("CRANE_B200_BENCH_EXTRA", "1") != "0":
        # BASELINE configs[2] beside the headline: Qwen3.5-0.8B, 4096-token prefill (chunkwise Gated-Delta-Net) + 512 decode tokens
        out["other_configs"] = {"qwen3_5_0_8b": other_config_summary("qwen3_5_0_8b")}
    print(json.dumps(out))
    model.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
