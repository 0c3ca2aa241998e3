#!/usr/bin/env python
"""Debug: the 512^3 replay-exact comparison (tests/test_gpu_baseline_configs.py) for one odometry mode, printing the mismatching voxels;
run it with and without KT_INT_SEQ_REPLAY=1 / KT_INT_PREP=0 to attribute a difference.  usage: replay_ab.py <odometry> <frames>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kintinuous_b200 as kb
from kintinuous_b200 import synth
from oracle import refbind
odo = int(sys.argv[1]); nframes = int(sys.argv[2])
V, ROWS, COLS = 512, 480, 640
ref = refbind.RefCuda(V)
cfg = kb.Config.default(vol=V, odometry=odo)
mine = kb.Tracker(cfg)
intr = np.array(synth.intrinsics(COLS, ROWS), np.float32)
ts = torch.zeros(V ** 3, dtype=torch.int16, device="cuda"); cs = torch.zeros(V ** 3 * 4, dtype=torch.uint8, device="cuda")
ref.init_volume(ts, cs)
fb = torch.zeros((ROWS, COLS), dtype=torch.int16, device="cuda")
vm = torch.zeros((3 * ROWS, COLS), dtype=torch.float32, device="cuda"); nm = torch.zeros_like(vm)
ds = torch.zeros((ROWS, COLS), dtype=torch.float32, device="cuda")
cur = [0, 0, 0]
first_bad = None
ops = kb.ops
t3 = torch.zeros(V ** 3, dtype=torch.int16, device="cuda"); c3 = torch.zeros(V ** 3 * 4, dtype=torch.uint8, device="cuda")     # OUR operator on the same arguments
ops.init_volume(t3, c3, V)
ds3 = torch.zeros((ROWS, COLS), dtype=torch.float32, device="cuda")
for k in range(nframes):
    d, c = synth.render(k)
    p = mine.process_frame(d, c, k)
    wa = np.array(p.voxel_wrap)
    dd = torch.from_numpy(d.view(np.int16)).cuda(); cc = torch.from_numpy(c).cuda()
    ref.bilateral(dd, fb, ROWS, COLS); ref.vmap(fb, vm, ROWS, COLS, intr); ref.nmap(vm, nm, ROWS, COLS)
    for axis in range(3):
        n = int(wa[axis]) - cur[axis]
        if n:
            ref.clear(axis, 1 if n < 0 else 0, ts, cs, cur[axis], cur[axis] + n)
            ops.clear_volume(axis, 1 if n < 0 else 0, t3, c3, V, cur[axis], cur[axis] + n); cur[axis] += n
    Rinv, tint, wint = mine.last_integrate()
    print("ARGS", k, " ".join(f"{int(x):08x}" for x in np.concatenate([Rinv.reshape(-1), tint]).astype(np.float32).view(np.uint32)), list(map(int, wint)), flush=True)
    ref.integrate(dd, ROWS, COLS, intr, [6.0] * 3, Rinv, tint, mine.trunc_dist, ts, cs, wint, cc, nm, 1, ds)
    ops.integrate(dd, ROWS, COLS, intr, [6.0] * 3, Rinv, tint, mine.trunc_dist, t3, c3, V, wint, cc, nm, 1, ds3)
    torch.cuda.synchronize()
    dsm = mine.download_map(6, 0); dsr = ds.cpu().numpy()
    nbad = int((dsm.view(np.uint32) != dsr.view(np.uint32)).sum())
    nbad3 = int((ds3.cpu().numpy().view(np.uint32) != dsr.view(np.uint32)).sum())
    if nbad or nbad3:
        print(f"frame {k}: scaled depth differs from the reference's at {nbad} (tracker) / {nbad3} (operator) pixels", flush=True)
    if first_bad is None and (k < 3 or k % 4 == 3 or k == nframes - 1):
        ta, ca = mine.export_volume()
        tr = ts.cpu().numpy().reshape(V, V, V); cr = cs.cpu().numpy().reshape(V, V, V, 4)
        bt = int((ta != tr).sum()); bc = int((ca != cr).any(-1).sum())
        bt3 = int((t3.cpu().numpy().reshape(V, V, V) != tr).sum()); bc3 = int((c3.cpu().numpy().reshape(V, V, V, 4) != cr).any(-1).sum())
        print(f"frame {k}: tracker vs reference replay: tsdf mismatches {bt}, colour mismatches {bc};  OUR OPERATOR vs reference replay: {bt3}, {bc3}", flush=True)
        if bt or bc:
            first_bad = k
            idx = np.argwhere((ca != cr).any(-1) | (ta != tr))[:6]
            for z, y, x in idx:
                print("  voxel", (int(z), int(y), int(x)), "mine", int(ta[z, y, x]), ca[z, y, x].tolist(), "replay", int(tr[z, y, x]), cr[z, y, x].tolist())
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("KT_"))
print(f"REPLAY_AB odometry={odo} frames={nframes} [{tag}] first frame with a mismatch: {first_bad}")
