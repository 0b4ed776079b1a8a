#!/usr/bin/env python
"""Per-kernel summary (launch count, min / median device time) of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys

hdr, agg = None, collections.OrderedDict()
for line in csv.reader(open(sys.argv[1])):
    if "Kernel Name" in line:
        hdr = line
        continue
    if hdr is None or len(line) < len(hdr) or not line[0].isdigit():
        continue
    name = re.sub(r"\(.*", "", line[hdr.index("Kernel Name")]).split("::")[-1]
    agg.setdefault(name, []).append(float(line[hdr.index("Metric Value")].replace(",", "")))
for k, v in agg.items():
    v = sorted(v)
    print(f"{k:40s} launches {len(v):3d}   min {v[0] / 1e3:8.1f} us   median {v[len(v) // 2] / 1e3:8.1f} us")
