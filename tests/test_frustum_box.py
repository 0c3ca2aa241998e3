"""kt_frustum.hpp (the box of storage tiles integrate_kernel is launched over) on the CPU: it must be a SUPERSET of the voxels the
reference's tsdf23 could update (cuda/tsdf_volume.cu:579-589: positive depth, rounded pixel inside the image) for any camera, and the
cyclic tile range must cover every storage coordinate of the logical range under any wrap."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libkt_frustum_host.so")
    src = os.path.join(ROOT, "tests", "cpp", "frustum_host.cpp")
    hdr = os.path.join(ROOT, "kintinuous_b200", "csrc", "kt_frustum.hpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "kintinuous_b200", "csrc"), "-o", so, src])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rot(rng, max_angle):
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    ang = rng.uniform(-max_angle, max_angle)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def test_box_contains_every_voxel_the_kernel_could_update(lib):
    rng = np.random.default_rng(7)
    V, size = 48, 6.0
    rows, cols = 120, 160
    k4 = np.array([132.0, 132.0, 80.0, 66.75], np.float32)
    cell = np.full(3, size / V, np.float32)
    idx = np.arange(V)
    gx, gy, gz = np.meshgrid(idx, idx, idx, indexing="ij")
    centres = (np.stack([gx, gy, gz], -1).reshape(-1, 3) + 0.5) * cell.astype(np.float64)
    n_nonempty = 0
    for trial in range(150):
        R = _rot(rng, np.pi)                                            # any orientation
        t = rng.uniform(-1.0, size + 1.0, 3).astype(np.float32)        # inside and outside the cube
        Rinv = np.linalg.inv(R).astype(np.float32)
        lo = np.zeros(3, np.int32); hi = np.zeros(3, np.int32)
        empty = lib.kth_frustum_box(_p(Rinv), _p(t), _p(k4), rows, cols, V, _p(cell), _p(lo), _p(hi))
        p = (centres - t.astype(np.float64)) @ Rinv.astype(np.float64).T
        with np.errstate(divide="ignore", invalid="ignore"):
            u = np.rint(k4[0] * p[:, 0] / p[:, 2] + k4[2]); v = np.rint(k4[1] * p[:, 1] / p[:, 2] + k4[3])
        can = (p[:, 2] > 0) & (u >= 0) & (u < cols) & (v >= 0) & (v < rows)
        if not can.any():
            continue
        assert not empty, (trial, t)
        n_nonempty += 1
        vox = np.stack([gx, gy, gz], -1).reshape(-1, 3)[can]
        assert (vox >= lo).all() and (vox <= hi).all(), (trial, vox.min(0), vox.max(0), lo, hi)
    assert n_nonempty > 60


def test_box_is_tight_for_a_centred_camera(lib):
    """the point of the exercise: a camera at the centre of the cube looking along +z reaches ~15 % of the columns"""
    V, size = 512, 6.0
    k4 = np.array([528.0, 528.0, 320.0, 267.0], np.float32)
    cell = np.full(3, size / V, np.float32)
    Rinv = np.eye(3, dtype=np.float32); t = np.full(3, 3.0, np.float32)
    lo = np.zeros(3, np.int32); hi = np.zeros(3, np.int32)
    assert lib.kth_frustum_box(_p(Rinv), _p(t), _p(k4), 480, 640, V, _p(cell), _p(lo), _p(hi)) == 0
    frac = np.prod((hi - lo + 1) / V)
    assert lo[2] >= V // 2 - 5 and hi[2] == V - 1 and frac < 0.2, (lo, hi, frac)


def test_cyclic_tile_range_covers_the_wrapped_range(lib):
    rng = np.random.default_rng(3)
    for V, tile in ((512, 32), (512, 8), (128, 32), (1024, 8)):
        for _ in range(300):
            lo = int(rng.integers(0, V)); hi = int(rng.integers(lo, V)); wrap = int(rng.integers(0, V))
            first = C.c_int(); n = C.c_int()
            lib.kth_cyclic_tile_range(lo, hi, wrap, V, tile, C.byref(first), C.byref(n))
            tiles = V // tile
            assert 1 <= n.value <= tiles and 0 <= first.value < tiles
            covered = {((first.value + i) % tiles) for i in range(n.value)}
            need = {((x + wrap) % V) // tile for x in range(lo, hi + 1)}
            assert need <= covered, (V, tile, lo, hi, wrap)
            assert len(covered) <= len(need) + 1 or n.value == tiles
