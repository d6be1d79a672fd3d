#!/usr/bin/env python
"""Per-source-line executed instructions of one ncu source page (see profiles/src_hotspots.py), as thread
instructions per matrix element.  Usage: stage_breakdown.py src.csv n_elements [min_per_value]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nval = float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
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
        out.append((int(r[smp]), int(r[ie]), fname, int(r[0]), r[1].strip()[:100]))
    except ValueError:
        pass
ts = sum(o[0] for o in out) or 1
ti = sum(o[1] for o in out) or 1
print(f"total warp instructions {ti}: {ti * 32 / nval:.1f} thread instructions per element; {ts} samples")
for s, i, f, n, src in out:
    if i * 32 / nval >= thr:
        print(f"{f}:{n:5d} {i * 32 / nval:6.1f}/elem  smp {100 * s / ts:4.1f}%  {src}")
