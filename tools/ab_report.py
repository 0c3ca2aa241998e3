#!/usr/bin/env python
"""A/B report on a GPU box: product kernels (kintinuous_b200, C ABI) vs the reference's own CUDA operators
(oracle/_ref/libkt_ref_<VOL>.so) on identical device buffers, then both trackers over a synthetic sequence.
Diagnostic tool (test infrastructure): prints one line per comparison and writes gpurun_out/ab_report_<VOL>.json.
usage: python tools/ab_report.py [--vol 256] [--frames 12] [--odometry 0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import kintinuous_b200 as kb  # noqa: E402
from kintinuous_b200 import synth  # noqa: E402
from oracle import refbind  # noqa: E402

REPORT = {}


def rec(name, **kw):
    REPORT[name] = kw
    print(f"[{name}] " + " ".join(f"{k}={v}" for k, v in kw.items()), flush=True)


def cmp_int(name, a, b):
    a = a.cpu().numpy().astype(np.int64); b = b.cpu().numpy().astype(np.int64)
    d = np.abs(a - b)
    rec(name, n=int(a.size), mismatch=int((d != 0).sum()), max_abs=int(d.max()) if d.size else 0)


def cmp_map(name, a, b, rows, cols):
    """SoA x3 float maps: compare NaN masks of the x plane and values where both valid."""
    a = a.cpu().numpy().reshape(3, rows, cols); b = b.cpu().numpy().reshape(3, rows, cols)
    na, nb = np.isnan(a[0]), np.isnan(b[0])
    both = ~na & ~nb
    d = np.abs(a[:, both] - b[:, both])
    bits = (a[:, both].view(np.uint32) != b[:, both].view(np.uint32)).sum()
    rec(name, valid=int(both.sum()), nan_mask_mismatch=int((na != nb).sum()), max_abs=float(d.max()) if d.size else 0.0, bit_mismatch=int(bits))


def canon(points):
    """Order-independent canonical form of a 32-byte point array (n x 4 u64, lexicographically sorted)."""
    arr = np.ascontiguousarray(points).view(np.uint64).reshape(len(points), 4)
    if len(arr) == 0:
        return arr
    return arr[np.lexsort(arr.T[::-1])]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vol", type=int, default=256)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--odometry", type=int, default=0)
    ap.add_argument("--skip-ops", action="store_true")
    ap.add_argument("--voxel-shift", type=int, default=14)
    args = ap.parse_args()
    V = args.vol
    dev = torch.device("cuda:0")
    ref = refbind.RefCuda(V)
    ops = kb.ops
    rows, cols = 480, 640
    fx, fy, cx, cy = synth.intrinsics(cols, rows)
    intr = np.array([fx, fy, cx, cy], np.float32)
    frames = [synth.render(k, cols, rows) for k in range(args.frames)]
    size = 6.0
    voxel = np.float32(size) / np.float32(V)
    trunc = float(max(np.float32(max(0.01, size / 100.0)), np.float32(2.1) * voxel))
    vs = [size, size, size]

    if not args.skip_ops:
        d0 = torch.from_numpy(frames[0][0].view(np.int16)).to(dev)
        rgb0 = torch.from_numpy(frames[0][1]).to(dev).contiguous()
        # --- bilateral / pyrdown
        fa = torch.zeros((rows, cols), dtype=torch.int16, device=dev); fb = torch.zeros_like(fa)
        ops.bilateral(d0, fa, rows, cols); ref.bilateral(d0, fb, rows, cols)
        cmp_int("bilateral", fa.view(torch.int16), fb.view(torch.int16))
        pyr_a, pyr_b = [fa], [fb]
        for l in range(1, 4):
            a = torch.zeros((rows >> l, cols >> l), dtype=torch.int16, device=dev); b = torch.zeros_like(a)
            ops.pyrdown(pyr_b[l - 1], a, rows >> (l - 1), cols >> (l - 1)); ref.pyrdown(pyr_b[l - 1], b, rows >> (l - 1), cols >> (l - 1))
            cmp_int(f"pyrdown_L{l}", a, b)
            pyr_a.append(a); pyr_b.append(b)
        # --- maps
        vm_b, nm_b = [], []
        for l in range(4):
            r, c = rows >> l, cols >> l
            k = intr / np.float32(1 << l)
            va = torch.zeros((3 * r, c), dtype=torch.float32, device=dev); na = torch.zeros_like(va)
            vb = torch.zeros_like(va); nb = torch.zeros_like(va)
            ops.create_maps(k, pyr_b[l], va, na, r, c)
            ref.vmap(pyr_b[l], vb, r, c, k); ref.nmap(vb, nb, r, c)
            cmp_map(f"vmap_L{l}", va, vb, r, c); cmp_map(f"nmap_L{l}", na, nb, r, c)
            v2 = torch.zeros_like(va); n2 = torch.zeros_like(va)
            ops.create_vmap(k, pyr_b[l], v2, r, c); ops.create_nmap(v2, n2, r, c)
            cmp_map(f"vmap_op_L{l}", v2, vb, r, c); cmp_map(f"nmap_op_L{l}", n2, nb, r, c)
            vm_b.append(vb); nm_b.append(nb)
        # --- transform maps (frame-0 model)
        R0 = np.eye(3, dtype=np.float32); t0 = np.array([3, 3, 3], np.float32)
        ang = 0.03; R1 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
        ga, gna = torch.zeros_like(vm_b[0]), torch.zeros_like(vm_b[0]); gb, gnb = torch.zeros_like(vm_b[0]), torch.zeros_like(vm_b[0])
        ops.transform_maps(vm_b[0], nm_b[0], R1, t0, ga, gna, rows, cols); ref.transform_maps(vm_b[0], nm_b[0], R1, t0, gb, gnb, rows, cols)
        cmp_map("transform_v", ga, gb, rows, cols); cmp_map("transform_n", gna, gnb, rows, cols)
        # --- resize
        for nm, fo, fr in (("resize_v", ops.resize_vmap, ref.resize_vmap), ("resize_n", ops.resize_nmap, ref.resize_nmap)):
            src = gb if nm == "resize_v" else gnb
            a = torch.zeros((3 * rows // 2, cols // 2), dtype=torch.float32, device=dev); b = torch.zeros_like(a)
            fo(src, a, rows, cols); fr(src, b, rows, cols)
            cmp_map(nm, a, b, rows // 2, cols // 2)
        # --- icp step: model = frame 0 maps in volume frame (identity rot), current = frame 3 maps
        mv, mn = torch.zeros_like(vm_b[0]), torch.zeros_like(vm_b[0])
        ref.transform_maps(vm_b[0], nm_b[0], R0, t0, mv, mn, rows, cols)
        d3 = torch.from_numpy(frames[min(3, args.frames - 1)][0].view(np.int16)).to(dev)
        f3 = torch.zeros_like(fb); ref.bilateral(d3, f3, rows, cols)
        cv, cn = torch.zeros_like(vm_b[0]), torch.zeros_like(vm_b[0])
        ref.vmap(f3, cv, rows, cols, intr); ref.nmap(cv, cn, rows, cols)
        Aa, ba, ra = ops.icp_step(R0, t0, cv, cn, R0, t0, intr, mv, mn, rows, cols)
        Ab, bb, rb = ref.icp_step(R0, t0, cv, cn, R0, t0, intr, mv, mn, rows, cols)
        rec("icp_step_L0", rel_A=float(np.abs(Aa - Ab).max() / np.abs(Ab).max()), rel_b=float(np.abs(ba - bb).max() / np.abs(bb).max()),
            inliers=(float(ra[1]), float(rb[1])), residual=(float(ra[0]), float(rb[0])))
        # --- integrate: two frames into fresh volumes, second with a wrapped volume
        ta = torch.zeros(V ** 3, dtype=torch.int16, device=dev); ca = torch.zeros(V ** 3 * 4, dtype=torch.uint8, device=dev)
        tb = torch.zeros_like(ta); cb = torch.zeros_like(ca)
        ops.init_volume(ta, ca, V); ref.init_volume(tb, cb)
        dsa = torch.zeros((rows, cols), dtype=torch.float32, device=dev); dsb = torch.zeros_like(dsa)
        for it, (wrap, R, t) in enumerate([((0, 0, 0), R0, t0), ((14, 3, 250), R1, t0 + np.array([0.02, -0.01, 0.03], np.float32))]):
            Rinv = np.linalg.inv(R.astype(np.float64)).astype(np.float32)
            t_a = time.time()
            ops.integrate(d0, rows, cols, intr, vs, Rinv, t, trunc, ta, ca, V, wrap, rgb0, nm_b[0], 1, dsa)
            t_b = time.time()
            ref.integrate(d0, rows, cols, intr, vs, Rinv, t, trunc, tb, cb, wrap, rgb0, nm_b[0], 1, dsb); torch.cuda.synchronize()
            t_c = time.time()
            dd = (dsa.view(torch.int32) != dsb.view(torch.int32)).sum().item()
            dt = (ta.to(torch.int32) - tb.to(torch.int32)).abs()
            dc = (ca.to(torch.int32) - cb.to(torch.int32)).abs().view(-1, 4)
            rec(f"integrate_{it}", scaled_bit_mismatch=int(dd), tsdf_nonzero=int((tb != 0).sum().item()), tsdf_mismatch=int((dt != 0).sum().item()),
                tsdf_max=int(dt.max().item()), tsdf_gt1=int((dt > 1).sum().item()), weight_mismatch=int((dc[:, 3] != 0).sum().item()),
                rgb_mismatch=int((dc[:, :3].sum(1) != 0).sum().item()), rgb_max=int(dc[:, :3].max().item()), ms_mine=round((t_b - t_a) * 1e3, 3), ms_ref=round((t_c - t_b) * 1e3, 3))
        # --- raycast from the reference volume
        for it, (wrap, R, t) in enumerate([((14, 3, 250), R1, t0 + np.array([0.02, -0.01, 0.03], np.float32))]):
            va = torch.zeros((3 * rows, cols), dtype=torch.float32, device=dev); na = torch.zeros_like(va); vb = torch.zeros_like(va); nb = torch.zeros_like(va)
            cca = torch.zeros((rows, cols, 4), dtype=torch.uint8, device=dev); ccb = torch.zeros_like(cca)
            t_a = time.time()
            ops.raycast(intr, R, t, trunc, vs, tb, V, va, na, rows, cols, wrap, cca, cb)
            t_b = time.time()
            ref.raycast(intr, R, t, trunc, vs, tb, vb, nb, rows, cols, wrap, ccb, cb); torch.cuda.synchronize()
            t_c = time.time()
            cmp_map(f"raycast_v_{it}", va, vb, rows, cols); cmp_map(f"raycast_n_{it}", na, nb, rows, cols)
            dcol = (cca.to(torch.int32) - ccb.to(torch.int32)).abs()
            rec(f"raycast_color_{it}", mismatch=int((dcol != 0).sum().item()), max=int(dcol.max().item()), ms_mine=round((t_b - t_a) * 1e3, 3), ms_ref=round((t_c - t_b) * 1e3, 3))
        # --- extract (thin slabs on every axis + full) and clear
        cap = 3 * rows * cols
        oa = torch.zeros(cap * 32, dtype=torch.uint8, device=dev); ob = torch.zeros_like(oa)
        wrap = (14, 3, 250); real = (14, 3, 250 - V)
        boxes = {"x+": (0, 17, 0, V, 0, V), "x-": (V - 16, V, 0, V, 0, V), "y+": (0, V, 0, 17, 0, V), "z+": (0, V, 0, V, 0, 17), "z-": (0, V, 0, V, V - 17, V - 1), "full": (0, V, 0, V, 0, V)}
        for nm, box in boxes.items():
            n_a = ops.extract_slice(tb, vs, V, oa, cap, wrap, cb, box, 1, real)
            n_b = ref.extract(tb, vs, ob, cap, wrap, cb, box, 1, real)
            pa = oa.cpu().numpy().view(refbind.POINT_DTYPE)[:n_a]; pb = ob.cpu().numpy().view(refbind.POINT_DTYPE)[:n_b]
            sa, sb = canon(pa), canon(pb)
            same = bool(sa.shape == sb.shape and (sa == sb).all())
            rec(f"extract_{nm}", n_mine=int(n_a), n_ref=int(n_b), multiset_equal=same)
        for axis in range(3):
            for back, (cur, n) in ((0, (14, 14)), (1, (-3, -14)), (0, (40, 16)), (1, (5, -16))):
                xa, ya = tb.clone(), cb.clone(); xb, yb = tb.clone(), cb.clone()
                ops.clear_volume(axis, back, xa, ya, V, cur, cur + n); ref.clear(axis, back, xb, yb, cur, cur + n); torch.cuda.synchronize()
                rec(f"clear_a{axis}_b{back}_n{n}", tsdf_mismatch=int((xa != xb).sum().item()), color_mismatch=int((ya != yb).sum().item()),
                    cleared=int(((xb == 0) & (tb != 0)).sum().item()))

    # ---------------- trackers over the sequence ----------------
    cfg = kb.Config.default(rows=rows, cols=cols, vol=V, odometry=args.odometry, voxel_shift=args.voxel_shift)
    mine = kb.Tracker(cfg)
    rt = ref.tracker(refbind.TrackerConfig.from_kt(cfg))
    poses = []
    t_mine = t_ref = 0.0
    for k, (d, c) in enumerate(frames):
        torch.cuda.synchronize(); t0_ = time.time()
        p = mine.process_frame(d, c, k)
        torch.cuda.synchronize(); t1_ = time.time()
        rt.process(d, c, k)
        torch.cuda.synchronize(); t2_ = time.time()
        if k > 0:
            t_mine += t1_ - t0_; t_ref += t2_ - t1_
        Ra, ta_, ga_, wa = p.as_tuple(); Rb, tb_, gb_, wb = rt.pose()
        dR = Ra.astype(np.float64) @ Rb.astype(np.float64).T
        ang = float(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
        gtR, gtt = synth.pose(k)
        poses.append(dict(frame=k, dt=float(np.abs(ta_ - tb_).max()), drot=ang, wrap_mine=wa.tolist(), wrap_ref=wb.tolist(),
                          err_gt_mine=float(np.abs((ga_ - gtt)).max()), err_gt_ref=float(np.abs((gb_ - gtt)).max())))
        tra, trb = mine.trace(), rt.trace()
        if k in (1, 2) and len(tra) and len(tra) == len(trb):
            rel = np.abs(tra[:, :42] - trb[:, :42]).max(1) / np.abs(trb[:, :42]).max(1)
            bad = np.flatnonzero(rel > 1e-4)
            rec(f"trace_frame{k}", iters=len(tra), max_rel=float(rel.max()), first_bad_iter=(int(bad[0]) if len(bad) else -1),
                counts_at_bad=((float(tra[bad[0], 43]), float(trb[bad[0], 43])) if len(bad) else None), inliers_first=(float(tra[0, 43]), float(trb[0, 43])), inliers_last=(float(tra[-1, 43]), float(trb[-1, 43])))
    rec("poses", max_dt=max(p["dt"] for p in poses), max_drot=max(p["drot"] for p in poses), last=poses[-1],
        err_gt_mine=max(p["err_gt_mine"] for p in poses), err_gt_ref=max(p["err_gt_ref"] for p in poses))
    REPORT["pose_list"] = poses
    ta, ca = mine.export_volume(); tb, cb = rt.export_volume()
    dt = np.abs(ta.astype(np.int32) - tb.astype(np.int32))
    dw = ca[..., 3] != cb[..., 3]
    drgb = np.abs(ca[..., :3].astype(np.int32) - cb[..., :3].astype(np.int32))
    rec("tracker_volume", nonzero=int((tb != 0).sum()), tsdf_mismatch=int((dt != 0).sum()), tsdf_gt1=int((dt > 1).sum()), tsdf_max=int(dt.max()),
        weight_mismatch=int(dw.sum()), rgb_mismatch=int((drgb.sum(-1) != 0).sum()), rgb_gt1=int((drgb.max(-1) > 1).sum()))
    rec("tracker_time", frames=len(frames) - 1, ms_per_frame_mine=round(1e3 * t_mine / max(1, len(frames) - 1), 3), ms_per_frame_ref=round(1e3 * t_ref / max(1, len(frames) - 1), 3))
    mine.finalise(); rt.finalise()
    pa, _, _ = mine.get_slice(mine.num_slices() - 1); pb, _, _ = rt.get_slice(rt.num_slices() - 1)
    rec("finalise", slices_mine=mine.num_slices(), slices_ref=rt.num_slices(), n_mine=len(pa), n_ref=len(pb))
    os.makedirs("gpurun_out", exist_ok=True)
    for i in range(min(mine.num_slices(), rt.num_slices())):
        qa, da, _ = mine.get_slice(i); qb, db, _ = rt.get_slice(i)
        ca_, cb_ = canon(qa), canon(qb)
        rec(f"slice_{i}", dim=(da, db), n=(len(qa), len(qb)), equal=bool(ca_.shape == cb_.shape and (ca_ == cb_).all()))
    with open(f"gpurun_out/ab_report_{V}_odo{args.odometry}_s{args.voxel_shift}.json", "w") as f:
        json.dump(REPORT, f, indent=1, default=str)


if __name__ == "__main__":
    main()
