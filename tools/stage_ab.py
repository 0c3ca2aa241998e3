#!/usr/bin/env python
"""A/B helper: mean CUDA-event stage times (ms) over a short run; tuning knobs come from the environment (KT_*).
usage: stage_ab.py [frames] [vol] [odometry]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kintinuous_b200 as kb
from kintinuous_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
vol = int(sys.argv[2]) if len(sys.argv) > 2 else 512
odo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
fr = [synth.render(k) for k in range(8)]
t = kb.Tracker(kb.Config.default(vol=vol, odometry=odo))
t.set_stage_timing(True)
acc = []
for i in range(n):
    k = i % 14; k = k if k < 8 else 14 - k
    p = t.process_frame(fr[k][0], fr[k][1], i)
    if i >= 6 and p.shifted == 0:
        acc.append(t.stage_ms())
m = np.mean(np.array(acc), axis=0)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("KT_"))
print(f"STAGES [{tag}] pyramid {m[0]:.4f} odometry {m[1]:.4f} shift {m[2]:.4f} integrate {m[3]:.4f} raycast {m[4]:.4f} total {m[5]:.4f}")
