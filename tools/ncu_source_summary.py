#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
warp instructions, thread instructions (=> average active lanes), stall samples.   usage: ncu_source_summary.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
H = rows[hi]
ci = {n: H.index(n) for n in ("Instructions Executed", "Thread Instructions Executed", "Predicated-On Thread Instructions Executed", "# Samples")}
src_col = 1
agg = collections.OrderedDict()
cur_file = ""
for r in rows[hi + 1:]:
    if len(r) < len(H) - 5:
        continue
    line, src = r[0], r[src_col]
    key = (line, src.strip()[:110])
    a = agg.setdefault(key, [0, 0, 0, 0])
    for k, n in enumerate(("Instructions Executed", "Thread Instructions Executed", "Predicated-On Thread Instructions Executed", "# Samples")):
        try:
            a[k] += float(r[ci[n]] or 0)
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()); tots = sum(a[3] for a in agg.values())
print(f"total warp inst {tot:.3e}, samples {tots:.0f}")
for (line, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    lanes = a[2] / a[0] if a[0] else 0
    print(f"{line:>5} {100 * a[0] / tot:5.1f}% inst  {100 * a[3] / max(1, tots):5.1f}% samples  lanes {lanes:4.1f}  | {src}")
