#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an ncu report.
usage: ncu_lines.py report.ncu-rep [min_pct]   (runs `ncu -i ... --page source --csv --print-source cuda,sass`)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
mn = float(sys.argv[2]) if len(sys.argv) > 2 else 1.2
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
src_cache = {}


def src_line(path, ln):
    if path not in src_cache:
        try:
            src_cache[path] = open(path.replace("/root/repo/", "")).read().split("\n")
        except OSError:
            src_cache[path] = []
    L = src_cache[path]
    return L[ln - 1].strip() if 0 < ln <= len(L) else ""


# sections: "File Path" row, "Function Name" row, header row, data rows
secs = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        path, func, hdr = rows[i][1], rows[i + 1][1], rows[i + 2]
        j = i + 3
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            data.append(rows[j])
            j += 1
        secs.append((path, func, hdr, data))
        i = j
    else:
        i += 1
by_func = {}
for path, func, hdr, data in secs:
    ie, sm, te = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    lsb = hdr.index("stall_long_sb") if "stall_long_sb" in hdr else None
    for r in data:
        if not r or r[0] == "" or len(r) <= te:
            continue
        try:
            ln, v, s, t = int(r[0]), float(r[ie]), int(r[sm]), float(r[te])
        except ValueError:
            continue
        l = int(r[lsb]) if lsb is not None and r[lsb].isdigit() else 0
        by_func.setdefault(func, []).append((path, ln, v, s, t, l))
for func, items in by_func.items():
    tot = sum(x[2] for x in items) or 1
    ts = sum(x[3] for x in items) or 1
    name = func.split("(")[0][-60:]
    print(f"===== {name}: inst {tot:.0f} samples {ts}")
    for path, ln, v, s, t, l in items:
        if 100 * v / tot >= mn or 100 * s / ts >= mn:
            print(f"{path.split('/')[-1][:14]:14s}:{ln:<4d} inst {100 * v / tot:5.1f}% samp {100 * s / ts:5.1f}% (long_sb {100 * l / ts:4.1f}%) thr {t / max(v, 1):4.1f} | {src_line(path, ln)[:90]}")
