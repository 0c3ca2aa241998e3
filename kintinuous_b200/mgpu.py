"""Host-side plumbing of the shared-volume mode (ONE stream, `world` GPUs, one process per GPU; `torch.distributed` for the control
plane only).

The data path has no collective library call.  The TSDF plane is replicated: the owner of a voxel stores its changed value into every
rank's replica from inside the integration kernel (NVLink P2P stores to CUDA-IPC mapped peer memory), so ray casting reads local HBM
only.  The colour / weight plane is sharded by storage z plane, block-cyclically (blocks of `block` planes dealt round-robin), which
spreads any viewing frustum evenly over the ranks.  The predicted-surface all-gather is P2P stores inside the ray-cast kernel,
synchronisation is a flag barrier in peer memory (kintinuous_b200/csrc/kt_tsdf.cu, xgpu_barrier_kernel).  `torch.distributed` (NCCL on
GPUs, gloo in the CPU tests) only carries the 64-byte IPC handles at start-up and the final timing reduction.
"""
from __future__ import annotations


def owner_of_plane(storage_z: int, world: int, block: int) -> int:
    """Rank that owns the colour / weight data (and performs the integration) of a storage z plane; invariant under volume shifting."""
    return (storage_z // block) % world


def local_plane(storage_z: int, world: int, block: int) -> int:
    """Index of an owned storage plane in its owner's local plane order."""
    return (storage_z // (block * world)) * block + storage_z % block


def owned_planes(rank: int, world: int, vol: int, block: int):
    """Storage z planes owned by `rank`, in local plane order (kt_volume_export_reference_layout / Tracker.export_owned order)."""
    n = vol // world
    return [((l // block) * world + rank) * block + l % block for l in range(n)]


def tile_rows(rank: int, world: int, rows: int, tile: int = 8):
    """Rows of 32x8 ray-cast tiles cast by `rank` (contiguous bands; every tile row belongs to exactly one rank)."""
    tiles = rows // tile
    return rank * tiles // world, (rank + 1) * tiles // world


def exchange(obj, group=None):
    """all-gather a small python object (the IPC handle) in rank order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


def connect(tracker, group=None):
    """Export this rank's arena handle, gather everybody's, map the peers.  Collective: every rank must call it."""
    import torch.distributed as dist
    handles = exchange(tracker.mgpu_arena_handle(), group)
    tracker.mgpu_connect(handles)
    dist.barrier(group)
    return handles


def icp_pixel_range(rank: int, world: int, cta: int, ctas: int, n_pixels: int):
    """Pixel range [begin, begin + count) of one CTA of icp_frame_kernel at a pyramid level with n_pixels pixels when the rows are split over
    the ranks of a shared volume (KT_MG_SPLIT_ICP; kt_icp.cu: GT = G * world CTAs, this one is number rank * G + cta, q pixels each, rounded
    up to the 16-byte TMA granule).  world = 1 is the single-GPU partition."""
    gt = ctas * world
    gci = rank * ctas + cta
    q = (((n_pixels + gt - 1) // gt) + 3) & ~3
    begin = min(n_pixels, gci * q)
    return begin, min(n_pixels, begin + q) - begin
