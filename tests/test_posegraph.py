"""kt_posegraph.hpp on the CPU: the .poses line of KintinuousTracker::outputPose (KintinuousTracker.cpp:199-218) and Eigen's
rotation -> quaternion conversion behind it, against scipy and against Python's own formatting of the same stream manipulators."""
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
    so = os.path.join(out, "libkt_posegraph_host.so")
    src = os.path.join(ROOT, "tests", "cpp", "posegraph_host.cpp")
    hdr = os.path.join(ROOT, "kintinuous_b200", "csrc", "kt_posegraph.hpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "kintinuous_b200", "csrc"), "-o", so, src])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_quaternion_matches_scipy_in_every_branch(lib):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(2)
    rots = [Rotation.random(random_state=int(s)) for s in rng.integers(0, 1 << 30, 400)]
    # rotations by ~pi about each axis exercise the three trace <= 0 branches
    for ax in np.eye(3):
        for a in (3.0, 3.1, np.pi, -3.05):
            rots.append(Rotation.from_rotvec(ax * a))
    branches = set()
    for r in rots:
        m = r.as_matrix().astype(np.float32)
        q = np.zeros(4, np.float32)
        lib.kth_quaternion(_p(np.ascontiguousarray(m)), _p(q))
        want = Rotation.from_matrix(m.astype(np.float64)).as_quat()                     # x, y, z, w
        if np.dot(want, q) < 0:
            want = -want
        assert np.abs(q - want).max() < 2e-6, (m, q, want)
        t = np.trace(m)
        branches.add(3 if t > 0 else int(np.argmax(np.diag(m))))
        if t > 0:
            assert q[3] > 0                                                              # Eigen's first branch: w = 0.5 sqrt(t + 1)
    assert branches == {0, 1, 2, 3}


def test_pose_line_format(lib):
    rng = np.random.default_rng(4)
    from scipy.spatial.transform import Rotation
    for _ in range(100):
        ts = int(rng.integers(0, 1 << 50))
        t = (rng.normal(size=3) * rng.choice([1e-3, 1.0, 100.0])).astype(np.float32)
        m = np.ascontiguousarray(Rotation.random(random_state=int(rng.integers(0, 1 << 30))).as_matrix().astype(np.float32))
        buf = C.create_string_buffer(256)
        n = lib.kth_pose_line(C.c_ulonglong(ts), _p(t), _p(m), buf, 256)
        line = buf.value.decode()
        assert n == len(line) and line.endswith("\n")
        f = line.split()
        assert len(f) == 8
        assert f[0] == "%.6f" % (ts / 1000000.0)                                         # setprecision(6) << fixed on the double
        q = np.zeros(4, np.float32); lib.kth_quaternion(_p(m), _p(q))
        want = ["%g" % float(v) for v in list(t) + list(q)]                            # operator<<(float): %g, 6 significant digits
        assert f[1:] == want, (f, want)
    buf = C.create_string_buffer(8)
    assert lib.kth_pose_line(C.c_ulonglong(1), _p(np.zeros(3, np.float32)), _p(np.eye(3, dtype=np.float32)), buf, 8) == -1
