#!/usr/bin/env python
"""Stage spans of the headline request (ViT + prefill of 454 positions + 256 decode steps) through crane_b200_prof_enable /
crane_b200_prof_report: device and host-submission time per stage, named as crane-core/src/ops/prof.rs names them.
    python tools/prefill_spans.py [--reps 5] [--precision split|bf16] [--out profiles/r02_spans.json]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import crane_b200  # noqa: E402
from crane_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default="")
args = ap.parse_args()
cfg = synth.QWEN3_VL_2B
m = crane_b200.Qwen3VLModel(cfg, device=0, max_seq_len=1024)
m.load_checkpoint(bench.synth_checkpoint_parallel(cfg, as_bits=True))
ids, pv, grid = bench.make_request(cfg)
for _ in range(2):
    m.clear_kv_cache()
    first = int(np.argmax(m.forward(ids, pv, [grid], 0)))
    m.decode_greedy(first, len(ids), 16)
plain = []
for _ in range(args.reps):
    m.clear_kv_cache()
    m.forward(ids, pv, [grid], 0)
    plain.append(m.last_timing()["prefill_ms"])
m.prof_enable(True)
for _ in range(args.reps):
    m.clear_kv_cache()
    first = int(np.argmax(m.forward(ids, pv, [grid], 0)))
    m.decode_greedy(first, len(ids), bench.N_DECODE)
rep = m.prof_report()
rep["unprofiled_prefill_ms"] = float(np.median(plain))
pre = rep["prefill"]
print(f"prefill: unprofiled {rep['unprofiled_prefill_ms']:.3f} ms; profiled: enqueue {pre['enqueue_ms']:.3f} wall {pre['wall_ms']:.3f} device {pre['device_ms']:.3f} ms")
for k, v in sorted(pre["spans"].items(), key=lambda kv: -kv[1]["device_ms"]):
    print(f"   {k:14s} device {v['device_ms']:8.3f} ms   host {v['host_ms']:7.3f} ms")
d = rep["decode"]
print(f"decode: {d['passes']} passes, device {d['device_ms']*1e3:.1f} us/pass, enqueue {d['enqueue_ms']*1e3:.2f} us/pass")
if args.out:
    json.dump(rep, open(args.out, "w"), indent=1)
m.close()
