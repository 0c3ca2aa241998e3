"""Host logic of the product's shifting volume (kintinuous_b200/csrc/kt_shift.hpp, used by kt_tracker.cu) on the CPU, against an
independent restatement of the reference lines it follows (KintinuousTracker.cpp:112, :581-596, :636-667, :675-831, :1075-1085;
TSDFVolume.cpp:96) and against invariants of the cyclic volume."""
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
    so = os.path.join(out, "libkt_shift_host.so")
    src = os.path.join(ROOT, "tests", "cpp", "shift_host.cpp")
    hdr = os.path.join(ROOT, "kintinuous_b200", "csrc", "kt_shift.hpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "kintinuous_b200", "csrc"), "-o", so, src])
    l = C.CDLL(so)
    l.kth_trunc_dist.restype = C.c_float
    l.kth_trunc_dist.argtypes = [C.c_float, C.c_float]
    l.kth_global_camera.restype = C.c_float
    l.kth_global_camera.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_float]
    l.kth_shift_steps.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_void_p]
    return l


def _i3(v=(0, 0, 0)):
    return (C.c_int * 3)(*v)


def test_trunc_distance_follows_the_reference(lib):
    for size, vol in ((6.0, 512), (6.0, 1024), (6.0, 2048), (3.0, 128), (0.5, 512), (20.0, 256)):
        voxel = np.float32(size) / np.float32(vol)
        want = max(max(np.float32(0.01), np.float32(size) / np.float32(100)), np.float32(2.1) * voxel)     # .cpp:112, TSDFVolume.cpp:96
        assert np.float32(lib.kth_trunc_dist(size, float(voxel))) == np.float32(want)
    assert abs(lib.kth_trunc_dist(6.0, 6.0 / 512) - 0.06) < 1e-7                                                # 0.06 m at 6 m


def test_nonnegative_wrap_alias_addresses_the_same_storage_plane(lib):
    rng = np.random.default_rng(0)
    for V in (128, 512, 1024):
        for _ in range(200):
            w = rng.integers(-5 * V, 5 * V, 3)
            out = _i3()
            lib.kth_vwrap_nonneg(_i3(w), V, out)
            o = np.array(list(out))
            for i in range(3):
                want = w[i] if w[i] >= 0 else V - ((-w[i]) % V)                     # .cpp:1075-1085 as written (V, not 0, for multiples of V)
                assert o[i] == want and o[i] >= 0
                assert (o[i] - w[i]) % V == 0                                        # congruent modulo V; it is NOT bounded: positive wraps pass through unchanged,
                # so every kernel wrapper reduces it with wrap_mod() before launch (kt_common.cuh; GPU test test_wrap_beyond_one_volume_length)


def test_shift_steps_clamp_and_floor(lib):
    rng = np.random.default_rng(1)
    voxel = np.float32(6.0 / 512)
    for thresh in (2, 14, 2**31 - 1):
        for _ in range(300):
            ct = (rng.standard_normal(3) * 0.3).astype(np.float32)
            tr = _i3()
            lib.kth_shift_steps(ct.ctypes.data_as(C.c_void_p), float(voxel), thresh, tr)
            for i in range(3):
                f = int(np.floor(np.float32(ct[i]) / voxel))                         # .cpp:642-667
                want = max(-thresh, f) if f < 0 else min(thresh, f)
                assert tr[i] == want


def test_shift_boxes(lib):
    V, overlap = 512, 2
    for thresh in (2, 14):
        for axis in range(3):
            for n in range(-thresh, thresh + 1):
                lo, hi = _i3(), _i3()
                d = lib.kth_shift_box(axis, n, thresh, overlap, V, lo, hi)
                lo, hi = list(lo), list(hi)
                for a in range(3):
                    if a != axis:
                        assert (lo[a], hi[a]) == (0, V)
                if n >= thresh:                                                      # .cpp:695 / :750 / :802
                    assert d == 1 and (lo[axis], hi[axis]) == (0, n + 1 + overlap)
                elif n <= -thresh:
                    if axis < 2:
                        assert d == -1 and (lo[axis], hi[axis]) == (V + (n - overlap), V)
                    else:                                                            # .cpp:805 (Q12): one plane lower for ZMinus
                        assert d == -1 and (lo[axis], hi[axis]) == (V + (n - overlap) - 1, V - 1)
                    assert hi[axis] - lo[axis] == -n + overlap
                else:
                    assert d == 0 and (lo[axis], hi[axis]) == (0, V)
    # parked: the threshold is INT_MAX and nothing ever shifts
    lo, hi = _i3(), _i3()
    assert lib.kth_shift_box(0, 10 ** 6, 2**31 - 1, overlap, V, lo, hi) == 0


def test_slice_dimension_codes(lib):
    codes = {(3, 0, 0): 0, (-3, 0, 0): 1, (0, 3, 0): 2, (0, -3, 0): 3, (0, 0, 3): 4, (0, 0, -3): 5}     # XPlus..ZMinus (CloudSlice.h:33-36)
    for vt, want in codes.items():
        assert lib.kth_slice_dimension(_i3(vt)) == want


def test_global_camera_is_invariant_under_a_shift(lib):
    """A shift by n voxels moves voxelWrap by n and the camera's in-volume translation by -n * voxel (.cpp:1168-1203): the global camera
    position must not jump (up to float rounding)."""
    size, V = 6.0, 512
    voxel = float(np.float32(size) / np.float32(V)); basis = size / 2
    rng = np.random.default_rng(2)
    for _ in range(200):
        wrap = int(rng.integers(-2000, 2000)); n = int(rng.integers(-14, 15)); t = float(rng.uniform(2.5, 3.5))
        before = lib.kth_global_camera(basis, size, wrap, voxel, t)
        after = lib.kth_global_camera(basis, size, wrap + n, voxel, float(np.float32(t) - np.float32(voxel * n)))
        assert abs(before - after) < 2e-5
        assert abs(before - ((basis - size / 2) + wrap * voxel + (t - basis))) < 1e-4
