#!/usr/bin/env python
"""Multi-GPU equivalence check (run under torchrun, one rank per GPU): ONE volume shared by `world` GPUs (replicated TSDF, block-cyclic
colour planes) must reproduce the single-GPU tracker bit for bit: poses, model maps, every rank's TSDF replica, the colour / weight planes
each rank owns, shift events, extracted slices as a multiset.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py --vol 256
"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import kintinuous_b200 as kb
from kintinuous_b200 import synth, mgpu

ap = argparse.ArgumentParser(); ap.add_argument("--vol", type=int, default=256); ap.add_argument("--frames", type=int, default=8); ap.add_argument("--voxel-shift", type=int, default=2)
ap.add_argument("--odometry", type=int, default=0)
ap.add_argument("--split-icp", action="store_true", help="KT_MG_SPLIT_ICP: pixel rows of the ICP split over the ranks, all-reduce fused into the kernel (poses then agree with the 1-GPU run to rounding, not bit for bit)")
args = ap.parse_args()
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
cfg = kb.Config.default(vol=args.vol, odometry=args.odometry, voxel_shift=args.voxel_shift, device=local, rank=rank, world=world)
if args.split_icp:
    os.environ["KT_MG_SPLIT_ICP"] = "1"           # read when the shared-volume context is created
trk = kb.Tracker(cfg)
os.environ.pop("KT_MG_SPLIT_ICP", None)
mgpu.connect(trk)
single = kb.Tracker(kb.Config.default(vol=args.vol, odometry=args.odometry, voxel_shift=args.voxel_shift, device=local))   # every rank also runs the 1-GPU tracker
frames = [synth.render(k) for k in range(args.frames)]
ok = True
t_m = t_s = 0.0
pose_max = 0.0
all_poses = []
for k, (d, c) in enumerate(frames):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    p = trk.process_frame(d, c, k)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    q = single.process_frame(d, c, k)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if k > 0: t_m += t1 - t0; t_s += t2 - t1
    all_poses.append(list(p.R) + list(p.t))
    pose_max = max(pose_max, float(np.abs(np.array(list(p.R) + list(p.t)) - np.array(list(q.R) + list(q.t))).max()))
    same = list(p.R) == list(q.R) and list(p.t) == list(q.t) and list(p.voxel_wrap) == list(q.voxel_wrap) and p.shifted == q.shifted
    if args.split_icp:
        same = pose_max <= 1e-5 and list(p.voxel_wrap) == list(q.voxel_wrap) and p.shifted == q.shifted
    if not same: ok = False; print(f"[rank {rank}] frame {k}: pose mismatch", np.abs(np.array(p.t) - np.array(q.t)).max(), list(p.voxel_wrap), list(q.voxel_wrap), flush=True)
torch.cuda.synchronize(); dist.barrier()
info = trk.mgpu_info()
if args.split_icp:
    # every rank must hold bit-identical poses (the totals are integer sums: identical on all ranks); against the 1-GPU run the poses agree to
    # rounding (a different grouping of the float partial sums), so the volumes are compared by the fraction of voxels within 1 LSB
    gathered_p = [None] * world
    dist.all_gather_object(gathered_p, all_poses)
    ranks_identical = all(g == gathered_p[0] for g in gathered_p)
    ta, _ = trk.export_owned(); tf, _ = single.export_volume()
    pl = mgpu.owned_planes(rank, world, args.vol, info["block"])
    frac = float((np.abs(ta.astype(np.int32) - tf[pl].astype(np.int32)) <= 1).mean())
    res = torch.tensor([int(ok), int(ranks_identical), int(frac >= 0.999)], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"MGPU_SPLIT_CHECK world={world} vol={args.vol} frames={args.frames} poses_within_1e-5={bool(res[0])} pose_max={pose_max:.2e} ranks_identical={bool(res[1])} "
              f"tsdf_within_1lsb={frac:.6f} ms_per_frame_shared={1e3 * t_m / (args.frames - 1):.3f} ms_per_frame_single={1e3 * t_s / (args.frames - 1):.3f}", flush=True)
    trk.close(); single.close()
    dist.destroy_process_group()
    sys.exit(0 if bool(res.min()) else 1)
planes = mgpu.owned_planes(rank, world, args.vol, info["block"])
ts, cs = trk.export_owned(); tf, cf = single.export_volume()
vol_ok = bool((ts == tf[planes]).all() and (cs == cf[planes]).all()) and bool((trk.export_tsdf_replica() == tf).all())
maps_ok = all(bool(np.array_equal(trk.download_map(w, l), single.download_map(w, l), equal_nan=True)) for w in (2, 3) for l in range(3))
trk.finalise(); single.finalise()
def canon(pts):
    a = np.ascontiguousarray(pts).view(np.uint64).reshape(len(pts), 4)
    return a[np.lexsort(a.T[::-1])] if len(a) else a
mine = [trk.get_slice(i)[0] for i in range(trk.num_slices())]
gathered = [None] * world
dist.all_gather_object(gathered, mine)
slices_ok = True
if rank == 0:
    for i in range(single.num_slices()):
        ref = canon(single.get_slice(i)[0])
        got = canon(np.concatenate([g[i] for g in gathered]))
        if ref.shape != got.shape or not (ref == got).all():
            slices_ok = False; print(f"slice {i}: ref {ref.shape} got {got.shape}", flush=True)
res = torch.tensor([int(ok), int(vol_ok), int(maps_ok), int(slices_ok)], device="cuda")
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"MGPU_CHECK world={world} vol={args.vol} frames={args.frames} poses_equal={bool(res[0])} slabs_equal={bool(res[1])} model_maps_equal={bool(res[2])} slices_equal={bool(res[3])} "
          f"ms_per_frame_zslab={1e3 * t_m / (args.frames - 1):.3f} ms_per_frame_single={1e3 * t_s / (args.frames - 1):.3f} slices={single.num_slices()}", flush=True)
trk.close(); single.close()
dist.destroy_process_group()
sys.exit(0 if bool(res.min()) else 1)
