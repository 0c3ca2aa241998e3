#!/usr/bin/env python
"""Per-SASS-instruction executed counts of an `ncu --page source --csv --print-source cuda,sass` export, in address order.
usage: ncu_sass_counts.py file.csv [min_million]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
H = rows[2]
ia = H.index("Address"); isrc = [i for i, h in enumerate(H) if h == "Source"]; ie = H.index("Instructions Executed"); it = H.index("Thread Instructions Executed"); isamp = H.index("# Samples")
seen = {}
for r in rows:
    if len(r) <= ie or not r[ia] or r[ia] in ("Address", "-"):
        continue
    try:
        seen[int(r[ia], 16)] = (r[isrc[1]], float(r[ie] or 0), float(r[it] or 0), float(r[isamp] or 0), r[0])
    except ValueError:
        pass
items = sorted(seen.items())
tot = sum(v[1] for k, v in items)
print("total warp inst", tot, "SASS instructions", len(items))
base = items[0][0]
for a, (s, e, t, sm, ln) in items:
    if e / 1e6 >= thr:
        print(f"{a - base:5x} L{ln:>4} {e / 1e6:8.3f}M lanes {t / e if e else 0:4.1f} smp {sm:5.0f} | {s[:100]}")
