#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
  launches CSV (gpu__time_duration per launch)  -> per-kernel totals / shares
  .ncu-rep (ncu --set full)                     -> selected raw metrics per captured launch"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path, out):
    rows = list(csv.reader(open(path, errors="ignore")))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[h]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("cb::", "")
        tot[n] += v
        cnt[n] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# {path}: {sum(cnt.values())} launches, {total:.1f} us total (ncu: cold-cache, serialised -- compare SHARES)\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for n, v in tot.most_common():
            f.write(f"{n},{cnt[n]},{v:.1f},{v / cnt[n]:.2f},{v / total:.4f}\n")


def report(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    with open(out, "w") as f:
        f.write(",".join(f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in idx) + "\n")
        for r in rows[2:]:
            f.write(",".join('"' + r[i] + '"' if "," in r[i] else r[i] for i in idx) + "\n")


def _raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            d[h] = (v, u)
        out.append(d)
    return out


def _bytes(d, key):
    v, u = d[key]
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def traffic(prefix, dst):
    """DRAM bytes of one decode step from the full captures prof_gemv_/prof_lmhead_/prof_attn_<tag>.ncu-rep:
    28 x (qkv + o + gate_up + down GEMV + attention) + lm_head, each = dram__bytes_read.sum + dram__bytes_write.sum of its launch."""
    import json
    tag = prefix
    gemv = _raw(f"gpurun_out/prof_gemv_{tag}.ncu-rep")
    head = _raw(f"gpurun_out/prof_lmhead_{tag}.ncu-rep")
    attn = _raw(f"gpurun_out/prof_attn_{tag}.ncu-rep")
    per = lambda d: _bytes(d, "dram__bytes_read.sum") + _bytes(d, "dram__bytes_write.sum")
    names = [d["Kernel Name"][0] for d in gemv]
    layer = sum(per(d) for d in gemv[:4])                      # four consecutive GEMVs of a layer (any rotation of qkv, o, gate_up, down)
    step = 28 * (layer + per(attn[0])) + per(head[0])
    json.dump({"decode_step_dram_bytes": step, "layer_gemv_dram_bytes": layer, "lm_head_dram_bytes": per(head[0]),
               "attention_dram_bytes": per(attn[0]), "gemv_kernels": names[:4],
               "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch ({tag})"}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    {"launches": launches, "report": report, "traffic": traffic}[kind](src, dst)
