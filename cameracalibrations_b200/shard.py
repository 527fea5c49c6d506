"""Sharding of the hot path over the GPUs of one box: one process per GPU.

Points, frames and views are independent: every rank takes a
contiguous range and no data-path collective exists.  The only exchange is the NCCL
all-reduce of the 21-double normal-equation block / the 4 error sums, done in
calibration.reproj_jtj / calibration.calculate_errors through torch.distributed.
"""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) of n items for `rank`; sizes differ by at most one and the
    ranges tile [0, n) in rank order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_frames(nframes: int, rank: int, world: int, ring: int = 4):
    """Frame range of `rank` split into ring-sized groups (a ring of device frame buffers
    is what a stream larger than HBM cycles through)."""
    lo, hi = shard_range(nframes, rank, world)
    return [(s, min(s + ring, hi)) for s in range(lo, hi, ring)]
