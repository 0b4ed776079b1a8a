#!/usr/bin/env python
"""Group an ncu `--metrics gpu__time_duration.sum --csv` launch list by (kernel, block size): count, mean/min/max us."""
import collections
import csv
import re
import sys

rows = collections.defaultdict(list)
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1000 if x["Metric Unit"] in ("ns", "nsecond") else v
        if len(sys.argv) < 3 or re.search(sys.argv[2], x["Kernel Name"]):
            rows[(re.sub(r"\(.*", "", x["Kernel Name"])[:60], x["Block Size"], x["Grid Size"])].append(v)
for k, v in sorted(rows.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[0]:60s} {k[1]:>14s} {k[2]:>14s} n={len(v):4d} mean={sum(v) / len(v):8.1f} min={min(v):8.1f} max={max(v):8.1f} total={sum(v):10.1f}")
