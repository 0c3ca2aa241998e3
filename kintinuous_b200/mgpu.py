"""Host-side plumbing of the z-slab mode (one process per GPU, `torch.distributed` for the control plane only).

The data path has no collective library call: slabs and model maps are reached through CUDA-IPC mapped peer memory over NVLink,
the predicted-surface all-gather is P2P stores inside the ray-cast kernel, synchronisation is a flag barrier in peer memory
(kintinuous_b200/csrc/kt_tsdf.cu, xgpu_barrier_kernel).  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) only carries
the 64-byte IPC handles at start-up and the final timing reduction.
"""
from __future__ import annotations


def slab_range(rank: int, world: int, vol: int):
    """Storage z planes owned by `rank`: contiguous, equal, invariant under volume shifting."""
    assert vol % world == 0
    s = vol // world
    return rank * s, (rank + 1) * s


def tile_rows(rank: int, world: int, rows: int, tile: int = 8):
    """Rows of 32x8 ray-cast tiles cast by `rank` (contiguous bands; every tile row belongs to exactly one rank)."""
    tiles = rows // tile
    return rank * tiles // world, (rank + 1) * tiles // world


def owner_of_plane(storage_z: int, world: int, vol: int) -> int:
    return storage_z // (vol // world)


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
