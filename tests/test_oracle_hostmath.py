"""The host-side restatements of the Eigen / OpenCV arithmetic the reference uses (oracle/kt_hostmath.hpp) against
independent implementations: numpy.linalg, scipy Rotation, and cv2.Rodrigues (python OpenCV is in this image)."""
import ctypes as C

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_ldlt6_solve_matches_numpy(cpu_oracle):
    rng = np.random.default_rng(20260922)
    for trial in range(50):
        J = rng.standard_normal((200, 6)) * np.array([1, 1, 1, 0.3, 0.3, 0.3])
        A = J.T @ J
        if trial % 5 == 0:
            A *= 1e4
        b = rng.standard_normal(6)
        x = np.zeros(6)
        cpu_oracle.lib.ktoracle_ldlt6_solve(_p(np.ascontiguousarray(A)), _p(b), _p(x))
        ref = np.linalg.solve(A, b)
        assert np.allclose(x, ref, rtol=1e-9, atol=1e-12)


def test_rodrigues_matches_cv2_and_scipy(cpu_oracle):
    cv2 = pytest.importorskip("cv2")
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(1)
    for scale in (1e-12, 1e-6, 1e-3, 0.1, 1.0, 3.0):
        for _ in range(10):
            r = rng.standard_normal(3) * scale
            R = np.zeros(9)
            cpu_oracle.lib.ktoracle_rodrigues(_p(r), _p(R))
            Rcv, _ = cv2.Rodrigues(r.reshape(3, 1))
            assert np.abs(R.reshape(3, 3) - Rcv).max() < 1e-14
            assert np.abs(R.reshape(3, 3) - Rotation.from_rotvec(r).as_matrix()).max() < 1e-12
    R = np.zeros(9)
    cpu_oracle.lib.ktoracle_rodrigues(_p(np.zeros(3)), _p(R))
    assert (R.reshape(3, 3) == np.eye(3)).all()


def test_mat3_inverse_matches_numpy(cpu_oracle):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(2)
    for _ in range(20):
        m = Rotation.from_rotvec(rng.standard_normal(3)).as_matrix().astype(np.float32)
        out = np.zeros(9, np.float32)
        cpu_oracle.lib.ktoracle_mat3_inverse(_p(np.ascontiguousarray(m)), _p(out))
        assert np.abs(out.reshape(3, 3) - np.linalg.inv(m.astype(np.float64))).max() < 5e-7
