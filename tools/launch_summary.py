#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): the kernels of the LAST `--step-marker`-delimited step
in launch order with their durations, and the per-kernel totals. Usage: launch_summary.py launches.csv [marker-kernel-substring]"""
import csv
import re
import sys
from collections import OrderedDict

rows, hdr = [], None
for r in csv.reader(open(sys.argv[1], errors="ignore")):
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
marker = sys.argv[2] if len(sys.argv) > 2 else "k_conum_regular"
names = [re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"])).replace("<unnamed>::", "") for r in rows]
dur = [float(r["Metric Value"].replace(",", "")) / 1e3 for r in rows]
starts = [i for i, n in enumerate(names) if marker in n]
if len(starts) < 2 and len(sys.argv) <= 2:  # lists captured before CoNum was split into two kernels
    marker = "k_conum_stage1"
    starts = [i for i, n in enumerate(names) if marker in n]
if len(starts) >= 2:
    a, b = starts[-2], starts[-1]
else:
    a, b = 0, len(names)
# a step is delimited by two consecutive markers (the marker kernel sits in the middle of the momentum segment: fine for totals)
tot = OrderedDict()
for n, d in zip(names[a:b], dur[a:b]):
    n = n[:70]
    t = tot.setdefault(n, [0, 0.0])
    t[0] += 1
    t[1] += d
print(f"{b - a} launches, {sum(dur[a:b]):.1f} us between the last two '{marker}' launches")
for n, (c, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{d:9.1f} us  x{c:<3d} {n}")
