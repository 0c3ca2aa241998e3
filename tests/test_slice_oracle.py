"""oracle/kt_slice_oracle.cpp (the checker of kt_slice.cu) pinned piecewise on the CPU.  PCL itself does not exist in this image, so the
restatement of pcl::VoxelGrid / pcl::NormalEstimation (PCL 1.7.2) is checked against independent statements of the same definitions:
numpy for the leaf assignment and centroids, numpy.linalg.eigh for the analytic eigen solver, scipy's cKDTree for the neighbour search,
an FP64 PCA for the normals."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests"))
from slice_cloud import make_cloud  # noqa: E402

LEAF = np.float32(6.0 / 512)


@pytest.fixture(scope="module")
def so():
    from oracle import refbind
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "libkt_slice_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libkt_slice_oracle.so"])
    return refbind.SliceOracle(), refbind


def test_weight_cull_and_voxel_grid_against_numpy(so):
    o, rb = so
    pts = make_cloud(point_dtype=rb.POINT_DTYPE)
    kept = o.weight_cull(pts, 8)
    assert len(kept) == int((pts["a"] >= 8).sum()) and (kept["a"] >= 8).all()
    assert np.array_equal(kept, pts[pts["a"] >= 8])                                    # input order preserved (CloudSliceProcessor.cpp:108-114)
    vg, min_b, div_b = o.voxel_grid(kept, float(LEAF))
    # independent statement of VoxelGrid::applyFilter: leaf = floor(p / leaf) - min_b, output sorted by leaf index, float centroid
    inv = np.float32(1.0) / LEAF
    xyz = np.stack([kept["x"], kept["y"], kept["z"]], -1)
    ijk = np.floor(xyz * inv).astype(np.int64)
    mb = np.floor(xyz.min(0) * inv).astype(np.int64); db = np.floor(xyz.max(0) * inv).astype(np.int64) - mb + 1
    assert (mb == min_b).all() and (db == div_b).all()
    idx = (ijk - mb) @ np.array([1, db[0], db[0] * db[1]])
    order = np.argsort(idx, kind="stable")
    uniq, start, cnt = np.unique(idx[order], return_index=True, return_counts=True)
    assert len(vg) == len(uniq)
    cen = np.add.reduceat(xyz[order].astype(np.float64), start) / cnt[:, None]
    got = np.stack([vg["x"], vg["y"], vg["z"]], -1)
    assert np.abs(got - cen).max() < 2e-6                                                # float sums of <= ~6 points a few metres from the origin
    for ch in ("r", "g", "b"):
        want = (np.add.reduceat(kept[ch][order].astype(np.float32), start) / cnt.astype(np.float32)).astype(np.int32)   # float mean, truncated
        assert np.array_equal(vg[ch].astype(np.int32), want), ch
    assert (vg["a"] == 0).all()                                                          # the packed rgb int has no alpha (voxel_grid.hpp)
    # every centroid lies in the leaf it stands for
    assert (np.floor(got * inv).astype(np.int64) - mb == np.stack([uniq % db[0], (uniq // db[0]) % db[1], uniq // (db[0] * db[1])], -1)).mean() > 0.9999


def test_eigen33_against_numpy(so):
    o, _ = so
    rng = np.random.default_rng(11)
    worst = 0.0
    for trial in range(300):
        a = rng.normal(size=(3, 20)) * np.array([[1.0], [0.6], [rng.uniform(0.01, 0.3)]])
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        p = (q @ a) * 0.03
        cov = np.cov(p, bias=True)
        ev, vec = o.eigen33(cov)
        w, v = np.linalg.eigh(cov)
        ang = np.arccos(min(1.0, abs(float(vec @ v[:, 0]))))
        worst = max(worst, ang)
        assert ang < 5e-3, (trial, ang, w)
        assert abs(ev - w[0]) <= 2e-3 * w[2] + 1e-12
    # a perfectly planar neighbourhood: the smallest root is clamped to 0 (computeRoots2 path)
    p = np.stack([rng.normal(size=20), rng.normal(size=20), np.zeros(20)]) * 0.03
    ev, vec = o.eigen33(np.cov(p, bias=True))
    assert abs(ev) < 1e-9 and abs(abs(vec[2]) - 1.0) < 1e-5


def test_normals_neighbour_search_is_exact_and_normals_follow_the_surface(so):
    from scipy.spatial import cKDTree
    o, rb = so
    pts = make_cloud(point_dtype=rb.POINT_DTYPE)
    vg, _, _ = o.voxel_grid(o.weight_cull(pts, 8), float(LEAF))
    out = o.normals(vg, 20, float(LEAF))
    xyz = np.stack([vg["x"], vg["y"], vg["z"]], -1).astype(np.float64)
    assert np.array_equal(np.stack([out["x"], out["y"], out["z"]], -1), np.stack([vg["x"], vg["y"], vg["z"]], -1))
    assert np.array_equal(out["r"], vg["r"]) and (out["_p0"] == 1.0).all()
    tree = cKDTree(xyz)
    dist, nn = tree.query(xyz, k=20)
    # FP64 PCA over the exact 20 nearest neighbours; the oracle's normals (PCL's float, single-pass covariance of raw coordinates) must
    # agree up to PCL's own cancellation noise, be unit length and point towards the origin
    nrm = np.stack([out["nx"], out["ny"], out["nz"]], -1).astype(np.float64)
    assert np.abs(np.linalg.norm(nrm, axis=1) - 1.0).max() < 1e-4
    assert ((-xyz * nrm).sum(1) >= -1e-6).all()                                          # flipNormalTowardsViewpoint(0, 0, 0)
    sel = np.random.default_rng(0).choice(len(xyz), 3000, replace=False)
    ang = []
    for i in sel:
        d = xyz[nn[i]] - xyz[nn[i]].mean(0)
        w, v = np.linalg.eigh(d.T @ d)
        if w[1] < 4 * w[0] + 1e-12:                                                      # no well-defined plane: skip
            continue
        ang.append(np.arccos(min(1.0, abs(float(v[:, 0] @ nrm[i])))))
    ang = np.array(ang)
    print(f"oracle (PCL float arithmetic) vs FP64 PCA: median {np.median(ang):.2e} rad, 99 % {np.quantile(ang, 0.99):.2e}, max {ang.max():.2e}")
    assert np.quantile(ang, 0.99) < 0.08 and np.median(ang) < 0.02
    # curvature = lambda0 / trace in [0, 1/3]
    assert (out["curvature"] >= 0).all() and (out["curvature"] <= 1.0 / 3 + 1e-3).all()
