"""Debug: extraction of the golden-fixture volume, product vs live reference vs golden file."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kintinuous_b200 as kb
from kintinuous_b200 import synth
from oracle import refbind
import test_gpu_ops as T
V = 256
g = np.load(os.path.join(ROOT, "tests", "golden", "ops_160x120.npz"))
rows, cols = 120, 160
intr = np.array(synth.intrinsics(cols, rows), np.float32)
d0, c0 = synth.render(0, cols, rows); d3, _ = synth.render(12, cols, rows)
ang = 0.03
R0 = np.eye(3, dtype=np.float32); t0 = np.array([3, 3, 3], np.float32)
R1 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32); t1 = t0 + np.array([0.02, -0.01, 0.03], np.float32)
e = dict(torch=torch, kb=kb, ops=kb.ops, ref=refbind.RefCuda(V), g=g, rows=rows, cols=cols, intr=intr, d0=d0, c0=c0, d3=d3, R0=R0, t0=t0, R1=R1, t1=t1, trunc=float(g["params"][4]))
ts, cs, ds0, wrap, vs = T._integrated_volume(e)
cap = 400000
real = (14, 3, 250 - V)
def canon(buf, n):
    a = buf.cpu().numpy()[: n * 32].view(np.uint64).reshape(n, 4)
    return a[np.lexsort(a.T[::-1])] if n else a
for rep in range(2):
  for name, box in {"zslab": (0, V, 0, V, 225, 242), "xplus": (0, 120, 0, V, 0, V), "yslab": (0, V, 180, 197, 0, V)}.items():
    oa = torch.zeros(cap * 32, dtype=torch.uint8, device="cuda"); ob = torch.zeros_like(oa)
    na = kb.ops.extract_slice(ts, vs, V, oa, cap, wrap, cs, box, 1, real)
    nb = e["ref"].extract(ts, vs, ob, cap, wrap, cs, box, 1, real)
    a, b, gg = canon(oa, na), canon(ob, nb), g[f"extract_{name}"]
    print(rep, name, "mine", na, "ref_live", nb, "golden", len(gg), "mine==ref_live", a.shape == b.shape and bool((a == b).all()), "mine==golden", a.shape == gg.shape and bool((a == gg).all()),
          "ref_live==golden", b.shape == gg.shape and bool((b == gg).all()), flush=True)
    if a.shape == gg.shape and not (a == gg).all():
        bad = np.flatnonzero((a != gg).any(1))
        print("  differing rows", len(bad), "first mine:", a[bad[0]].view(np.float32)[:3], hex(int(a[bad[0]][2])), " golden:", gg[bad[0]].view(np.float32)[:3], hex(int(gg[bad[0]][2])))
