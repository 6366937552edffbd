#!/usr/bin/env python
"""Per-opcode and hottest-instruction breakdown from `ncu -i X.ncu-rep --page source --csv > file`.
usage: ncu_hot.py file.csv section_index [unit_count] [min_count]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
k = int(sys.argv[2])
unit = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
mn = float(sys.argv[4]) if len(sys.argv) > 4 else 0.25
sec = rows[starts[k]:starts[k + 1]]
print(sec[0][1][-70:])
hdr, data = sec[1], sec[2:]
ie, src, samp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
te = hdr.index("Thread Instructions Executed")
data = [r for r in data if len(r) > ie and r[ie].replace('.', '').isdigit()]
tot = sum(float(r[ie]) for r in data)
print("rows", len(data), "total inst", tot, "per unit", round(tot / unit, 1), "samples", sum(int(r[samp]) for r in data))
ops = collections.Counter()
for r in data:
    t = r[src].split()
    if not t:
        continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    ops[op.split('.')[0]] += float(r[ie])
print(" ".join(f"{k}:{v / unit:.1f}" for k, v in ops.most_common(30)))
for r in data:
    if float(r[ie]) / unit >= mn:
        print(f"{r[src].strip()[:100]:100s} {float(r[ie]) / unit:7.2f} thr={float(r[te]) / max(float(r[ie]), 1):5.1f} s={r[samp]}")
