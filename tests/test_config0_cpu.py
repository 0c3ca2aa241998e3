"""BASELINE.json configs[0]: "single synthetic 640x480 depth frame, 128^3 TSDF, 1 ICP iter -- naive CPU loop (plumbing / correctness, no
GPU)".  The CPU loop of the same equations is the oracle port (oracle/kt_oracle_cpu.cpp; Eigen is not installable here, numpy stands in for
the 6x6 solve, cv2 for Rodrigues).  No golden vector exists at this size, so the checks are the properties one Gauss-Newton step and one
fusion must have on the analytic scene: the normal matrix is symmetric positive definite, the step moves the pose towards the generator's
ground truth and lowers the point-to-plane error, and the surface predicted from the fused 128^3 volume reproduces the input depth to about
a voxel (4.7 cm)."""
import ctypes as C

import numpy as np


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1))


def test_config0_one_icp_iteration_and_one_fusion_on_the_cpu(cpu_oracle):
    import cv2
    from kintinuous_b200 import synth
    lib = cpu_oracle.lib
    lib.ktoracle_set_threads(min(8, lib.ktoracle_hardware_threads() or 1))
    rows, cols, V, size = 480, 640, 128, 6.0
    intr = np.array(synth.intrinsics(cols, rows), np.float32)
    dist_thres = C.c_float(0.10); angle_thres = C.c_float(float(np.sin(np.float32(20.0) * np.float32(3.14159254) / np.float32(180.0))))

    def maps(depth):
        fb = np.zeros((rows, cols), np.uint16); lib.ktoracle_bilateral(_p(depth), _p(fb), rows, cols)
        vm = np.zeros((3 * rows, cols), np.float32); nm = np.zeros_like(vm)
        lib.ktoracle_vmap(_p(fb), _p(vm), rows, cols, _p(intr)); lib.ktoracle_nmap(_p(vm), _p(nm), rows, cols)
        return vm, nm

    # frame 0 defines the model (KintinuousTracker.cpp:481-557: the first frame's maps, transformed by the initial pose, ARE the model)
    k1 = 3
    d0, c0 = synth.render(0); d1, c1 = synth.render(k1)
    R0 = np.eye(3, dtype=np.float32); t0 = np.array([size / 2] * 3, np.float32)
    v0, n0 = maps(d0)
    mv = np.zeros_like(v0); mn = np.zeros_like(v0)
    lib.ktoracle_transform_maps(_p(v0), _p(n0), _p(_f(R0)), _p(_f(t0)), _p(mv), _p(mn), rows, cols)
    v1, n1 = maps(d1)

    def step(Rc, tc):
        A = np.zeros(36, np.float32); b = np.zeros(6, np.float32); res = np.zeros(2, np.float32)
        lib.ktoracle_icp_step(_p(_f(Rc)), _p(_f(tc)), _p(v1), _p(n1), _p(_f(R0)), _p(_f(t0)), _p(intr), _p(mv), _p(mn), rows, cols, dist_thres, angle_thres, _p(A), _p(b), _p(res))
        return A.reshape(6, 6).astype(np.float64), b.astype(np.float64), res

    # ---- ONE Gauss-Newton iteration at full resolution (ICPOdometry.cpp:127-178) ----
    A, b, res0 = step(R0, t0)
    assert np.allclose(A, A.T) and np.linalg.eigvalsh(A).min() > 0                     # 6 DoF constrained by room + sphere + cube
    valid = int((d1 > 0).sum())
    assert res0[1] > 0.5 * valid                                                        # most pixels find a model point within 10 cm / 20 degrees
    x = np.linalg.solve(A, b)
    Rinc, _ = cv2.Rodrigues(x[3:6])
    M = np.eye(4); M[:3, :3] = Rinc; M[:3, 3] = x[:3]                                    # resultRt = Rt * I
    Tprev = np.eye(4); Tprev[:3, :3] = R0; Tprev[:3, 3] = t0
    Tc = Tprev @ np.linalg.inv(M)                                                        # T_curr = T_prev * resultRt^-1
    Rg, tg = synth.pose(k1)                                                              # ground truth, world frame = first camera
    err0 = np.linalg.norm(tg)                                                            # error of the initial guess (identity)
    err1 = np.linalg.norm(Tc[:3, 3] - t0 - tg)
    ang0 = np.linalg.norm(cv2.Rodrigues(Rg)[0]); ang1 = np.linalg.norm(cv2.Rodrigues(Rg.T @ Tc[:3, :3])[0])
    assert err1 < 0.5 * err0 and ang1 < 0.5 * ang0, (err0, err1, ang0, ang1)             # one linearised step removes most of a 3 cm / 0.6 degree motion
    _, _, res1 = step(Tc[:3, :3].astype(np.float32), Tc[:3, 3].astype(np.float32))
    assert res1[0] / res1[1] < 0.5 * res0[0] / res0[1]                                   # mean squared point-to-plane error per inlier drops

    # ---- ONE fusion of each frame into the 128^3 volume, then the predicted surface (tsdf_volume.cu:541-640, ray_caster.cu:298-425) ----
    vs = _f([size] * 3); trunc = max(0.06, 2.1 * size / V)
    tsdf = np.zeros(V ** 3, np.int16); color = np.zeros(V ** 3 * 4, np.uint8); wrap = np.zeros(3, np.int32); ds = np.zeros((rows, cols), np.float32)
    lib.ktoracle_integrate(_p(d0), rows, cols, _p(intr), _p(vs), _p(_f(np.linalg.inv(R0))), _p(_f(t0)), C.c_float(trunc), _p(tsdf), _p(color), V, _p(wrap),
                           _p(np.ascontiguousarray(c0)), _p(n0), 1, _p(ds))
    Rc = Tc[:3, :3].astype(np.float32); tc = Tc[:3, 3].astype(np.float32)
    lib.ktoracle_integrate(_p(d1), rows, cols, _p(intr), _p(vs), _p(_f(np.linalg.inv(Rc.astype(np.float64)))), _p(_f(tc)), C.c_float(trunc), _p(tsdf), _p(color), V, _p(wrap),
                           _p(np.ascontiguousarray(c1)), _p(n1), 1, _p(ds))
    w = color.reshape(-1, 4)[:, 3]
    assert int((w == 2).sum()) > 30000 and int(w.max()) == 2                             # the overlap of the two frusta was fused twice
    pv = np.zeros((3 * rows, cols), np.float32); pn = np.zeros_like(pv); pc = np.zeros((rows, cols, 4), np.uint8)
    lib.ktoracle_raycast(_p(intr), _p(_f(Rc)), _p(_f(tc)), C.c_float(trunc), _p(vs), _p(tsdf), V, _p(pv), _p(pn), rows, cols, _p(wrap), _p(pc), _p(color))
    pv = pv.reshape(3, rows, cols)
    hit = ~np.isnan(pv[0]) & (d1 > 0)
    assert hit.mean() > 0.8
    cam = (pv[:, hit].T - tc) @ Rc                                                       # volume frame -> camera frame: R^T (p - t)
    dz = np.abs(cam[:, 2] - d1[hit] / 1000.0)
    assert np.median(dz) < 0.5 * size / V and np.quantile(dz, 0.9) < 1.5 * size / V, (np.median(dz), np.quantile(dz, 0.9))
