#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel name."""
import csv, sys, collections
lines = open(sys.argv[1]).read().splitlines()
i = [n for n, l in enumerate(lines) if l.startswith('"ID"')][0]
agg = collections.OrderedDict()
for r in csv.DictReader(lines[i:]):
    k = r['Kernel Name'].split('(')[0]
    agg.setdefault(k, []).append(float(r['Metric Value']) / 1e6)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:46s} n={len(v):3d} avg={sum(v)/len(v):8.3f} ms  min={min(v):8.3f}  share={100*sum(v)/tot:5.1f}%")
