#!/usr/bin/env python
"""Debug: per-iteration phase breakdown (SM cycles) of the whole-frame ICP kernel, from CTA 0."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kintinuous_b200 as kb
from kintinuous_b200 import synth
cfg = kb.Config.default(vol=512)
t = kb.Tracker(cfg)
t.set_stage_timing(True)
for k in range(6):
    d, c = synth.render(k); t.process_frame(d, c, k)
buf = np.zeros(512, np.int64)
kb.load().kt_debug_icp_profile(t.h, buf.ctypes.data_as(C.c_void_p))
full = buf.reshape(64, 8)[:19]
st = full[:, :5]
names = ["main+cta_reduce", "barrier", "sum_partials", "solve"]
d = np.diff(st, axis=1)
print("iter  " + "  ".join(f"{n:>16s}" for n in names) + "   total   gap_to_next")
for i in range(19):
    gap = st[i + 1, 0] - st[i, 4] if i < 18 else 0
    print(f"{i:3d}   " + "  ".join(f"{int(x):16d}" for x in d[i]) + f"  {int(st[i,4]-st[i,0]):7d}  {int(gap):6d}")
sol = np.stack([full[:, 5] - full[:, 3], full[:, 6] - full[:, 5], full[:, 7] - full[:, 6], full[:, 4] - full[:, 7]], 1)
print("solve split (unpack, ldlt+subst, rodrigues, matmul+pose), mean cycles:", sol.mean(0).astype(int).tolist())
print("sum cycles", int(st[18, 4] - st[0, 0]), "stage_ms", t.stage_ms())
