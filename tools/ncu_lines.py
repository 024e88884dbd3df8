#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`)
per CUDA source line: instructions executed, stall samples.  Usage: ncu_lines.py file.csv [top_n] [kernel_regex]"""
import csv, sys, re, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kre = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
rows = csv.reader(open(path, newline=''))
cur_file = None; cur_fn = None; hdr = None; cur_line = None; active = True
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1]; active = (kre is None or kre.search(cur_fn) is not None); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not active: continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0]), r[1].strip()[:110]); continue
    if len(r) < 9 or r[2] == "...": continue
    try:
        inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    a = agg.setdefault(cur_line, [0, 0, 0]); a[0] += inst; a[1] += samp; a[2] += 1
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print(f"total inst {ti:.3e} samples {ts}")
print("--- by instructions")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/ti:5.1f}% inst {100*a[1]/ts:5.1f}% smp sass={a[2]:4d} {k[0]}:{k[1]}: {k[2]}")
print("--- by samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*a[0]/ti:5.1f}% inst {100*a[1]/ts:5.1f}% smp sass={a[2]:4d} {k[0]}:{k[1]}: {k[2]}")
