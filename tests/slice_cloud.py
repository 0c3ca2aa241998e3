"""Synthetic stand-in for an extracted cloud slice (test input only): points on a bumpy surface sampled on the TSDF lattice the way
extractCloudSlice emits them -- up to three edge-interpolated points per surface voxel -- a few metres from the origin, with colours
and per-point weights (alpha) on both sides of the cull threshold."""
import numpy as np


def make_cloud(n_side=160, cell=6.0 / 512, seed=5, offset=(1.7, -0.9, 2.3), point_dtype=None):
    rng = np.random.default_rng(seed)
    ix, iy = np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij")
    x = (ix + 0.5) * cell; y = (iy + 0.5) * cell
    h = 0.25 + 0.06 * np.sin(9.0 * x) * np.cos(7.0 * y) + 0.4 * x                      # height field z = h(x, y): ~ one crossing per column
    pts = []
    for jitter_axis in range(3):                                                        # x-, y- and z-edge crossings of the same surface
        px = x + (rng.uniform(0, cell, x.shape) if jitter_axis == 0 else 0.0)
        py = y + (rng.uniform(0, cell, x.shape) if jitter_axis == 1 else 0.0)
        pz = 0.25 + 0.06 * np.sin(9.0 * px) * np.cos(7.0 * py) + 0.4 * px
        keep = rng.uniform(size=x.shape) < (0.9 if jitter_axis == 2 else 0.45)
        pts.append(np.stack([px[keep], py[keep], pz[keep]], -1))
    p = np.concatenate(pts).astype(np.float32) + np.asarray(offset, np.float32)
    # a handful of isolated points far from the sheet (the kNN search has to widen for them)
    far = (rng.uniform(-1, 1, (12, 3)) * 0.4 + np.array([0.9, 0.9, 1.6])).astype(np.float32) + np.asarray(offset, np.float32)
    p = np.concatenate([p, far])
    out = np.zeros(len(p), point_dtype)
    out["x"], out["y"], out["z"] = p[:, 0], p[:, 1], p[:, 2]
    out["r"] = rng.integers(0, 256, len(p)); out["g"] = rng.integers(0, 256, len(p)); out["b"] = rng.integers(0, 256, len(p))
    out["a"] = rng.integers(1, 40, len(p))
    return out[rng.permutation(len(out))]
