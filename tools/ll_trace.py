#!/usr/bin/env python
"""Print the event trace the persistent decode kernel writes with CRANE_B200_LL_TRACE=<file> (decode_ll.cu `trace`):
per (CTA, warp) the time of every event of one decode step on the common %globaltimer clock."""
import sys

import numpy as np

CAP = 4096
NAMES = {1: "phase_enter", 2: "x_loaded", 3: "slot_landed", 4: "consume_done", 5: "cta_barrier", 6: "epilogue_done", 7: "attn_enter",
         8: "q_ready", 9: "attn_done", 10: "merge_done", 11: "token_done"}
raw = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(6, CAP, 2)
maxp = int(sys.argv[2]) if len(sys.argv) > 2 else 12
t0 = min(int(r[0, 1]) for r in raw if r[0, 1])
for wi, r in enumerate(raw):
    cta = (0, 73, 140)[wi // 2]
    warp = (0, 9)[wi % 2]
    print(f"--- CTA {cta} warp {warp}")
    last = None
    nslot = 0
    for ev, t in r:
        ev, t = int(ev), int(t)
        if t == 0:
            break
        e, p = ev & 0xff, ev >> 8
        if p >= maxp and p < 112:
            continue
        if e == 3:
            nslot += 1
            if last is not None and t - last < 300:
                last = t
                continue
        dt = 0 if last is None else t - last
        print(f"  p{p:3d} {NAMES.get(e, e):14s} t={(t - t0) / 1e3:9.2f} us  (+{dt / 1e3:6.2f})" + (f"  slots so far {nslot}" if e == 3 else ""))
        last = t
