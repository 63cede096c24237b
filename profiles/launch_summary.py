#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the launch count and total / mean duration.
usage: launch_summary.py launches.csv [last_n_launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
out = []
for r in rows[hdr + 2:]:
    if len(r) > vi:
        try:
            out.append((int(r[ii]), r[ki], float(r[vi].replace(",", "")) / 1e6))
        except ValueError:
            pass
if len(sys.argv) > 2:
    out = out[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for _, k, ms in out:
    k = k[:90]
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + ms)
total = sum(t for _, t in agg.values())
print(f"{len(out)} launches, {total:.3f} ms")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.3f} ms {100 * t / total:5.1f}%  x{c:<4d} {k}")
