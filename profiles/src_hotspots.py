#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
share of warp-stall samples and of executed warp instructions.
Usage: src_hotspots.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fname = ""
h = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        smp = h.index("# Samples")
        ie = h.index("Instructions Executed")
        continue
    if h is None or len(r) <= ie or not r[0].isdigit():
        continue
    try:
        out.append((int(r[smp]), int(r[ie]), f"{fname}:{r[0]}", r[1].strip()[:110]))
    except ValueError:
        pass
tot_s = sum(o[0] for o in out) or 1
tot_i = sum(o[1] for o in out) or 1
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
out.sort(reverse=True)
for s, i, ln, src in out[:top]:
    print(f"{100 * s / tot_s:5.1f}% smp {100 * i / tot_i:5.1f}% inst  {ln}: {src}")
