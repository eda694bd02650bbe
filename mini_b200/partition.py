"""Swizzled-cyclic 1D vertex partition -- the NumPy mirror of b200::Partition (include/b200/partition.cuh).

owner(v) = (v & (P-1)) ^ swizzle(v >> log_p),  row(v) = v >> log_p,  swizzle(r) = top log_p bits of
(r * 0x9E3779B1 mod 2^32).  Host-side helpers only (tests, verification, label gathering); the kernels use
the C++ struct."""
from __future__ import annotations

import numpy as np

GOLDEN = 0x9E3779B1


def log2_ranks(world: int) -> int:
    assert world >= 1 and world & (world - 1) == 0, "the number of ranks must be a power of two"
    return world.bit_length() - 1


def swizzle(rows, world: int):
    k = log2_ranks(world)
    r = np.asarray(rows).astype(np.uint64)
    if k == 0:
        return np.zeros_like(r)
    return ((r * np.uint64(GOLDEN)) & np.uint64(0xFFFFFFFF)) >> np.uint64(32 - k)


def row(v, world: int):
    return np.asarray(v).astype(np.uint64) >> np.uint64(log2_ranks(world))


def owner(v, world: int):
    vv = np.asarray(v).astype(np.uint64)
    return (vv & np.uint64(world - 1)) ^ swizzle(vv >> np.uint64(log2_ranks(world)), world)


def bit(v, world: int, n_local: int):
    """Rank-major bitmap index of vertex v."""
    return owner(v, world) * np.uint64(n_local) + row(v, world)


def global_ids(rank: int, world: int, n_local: int):
    """Global vertex id of every local row of `rank` (int64 array of length n_local)."""
    r = np.arange(n_local, dtype=np.uint64)
    return ((r << np.uint64(log2_ranks(world))) | (np.uint64(rank) ^ swizzle(r, world))).astype(np.int64)
