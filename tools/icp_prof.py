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
st = full[:, :4]
names = ["main+cta_reduce", "exchange", "solve"]
d = np.diff(st, axis=1)
print("iter  " + "  ".join(f"{n:>16s}" for n in names) + "   total   gap_to_next")
for i in range(19):
    gap = st[i + 1, 0] - st[i, 3] if i < 18 else 0
    print(f"{i:3d}   " + "  ".join(f"{int(x):16d}" for x in d[i]) + f"  {int(st[i,3]-st[i,0]):7d}  {int(gap):6d}")
print("sum cycles", int(st[18, 3] - st[0, 0]), "stage_ms", t.stage_ms(), "icp_kernel_ms", t.icp_kernel_ms())
