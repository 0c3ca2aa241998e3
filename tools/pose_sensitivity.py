#!/usr/bin/env python
"""How far does the REFERENCE's own trajectory move under a 1-LSB change of its input?  (test infrastructure: uses oracle/_ref)

Runs, on the same synthetic 640x480 stream into 512^3 (BASELINE configs[1] / configs[2]):
  mine   the product tracker
  ref    the reference's CUDA kernels behind the restated host loop
  ref'   the same reference, with ONE depth pixel of frame 1 raised by 1 mm (the smallest possible input change)
and prints per frame |t_mine - t_ref| and |t_ref' - t_ref| per axis.  The closed loop (pose -> fused volume -> predicted surface ->
next pose) amplifies any difference along the weakly constrained direction of the scene; the product-vs-reference deviation has to
be read against the reference's own sensitivity.
usage: pose_sensitivity.py [odometry] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kintinuous_b200 as kb
from kintinuous_b200 import synth
from oracle import refbind

odo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 72
V = 512
cfg = kb.Config.default(vol=V, odometry=odo)
ref = refbind.RefCuda(V)
mine = kb.Tracker(cfg)
ra = ref.tracker(refbind.TrackerConfig.from_kt(cfg))
rb = ref.tracker(refbind.TrackerConfig.from_kt(cfg))
print("frame  |mine-ref| x y z        |ref'-ref| x y z      shifts")
for k in range(n):
    d, c = synth.render(k)
    d2 = d
    if k == 1:
        d2 = d.copy(); d2[240, 320] += 1
    p = mine.process_frame(d, c, k); ra.process(d, c, k); rb.process(d2, c, k)
    _, ta, ga, wa = p.as_tuple(); _, tb, gb, wb = ra.pose(); _, tc, gc, wc = rb.pose()
    e1 = np.abs(ga - gb); e2 = np.abs(gc - gb)
    print(f"{k:3d}  {e1[0]:.2e} {e1[1]:.2e} {e1[2]:.2e}   {e2[0]:.2e} {e2[1]:.2e} {e2[2]:.2e}   {wa.tolist()} {wb.tolist()} {wc.tolist()}")
