"""The swizzled-cyclic vertex partition: the NumPy mirror (mini_b200/partition.py) against the C ABI's
host arithmetic (b200_partition_*), bijectivity, and the arc balance it exists for (CPU only)."""
import ctypes as C

import numpy as np
import pytest

import mini_b200
from mini_b200 import partition as PT


@pytest.fixture(scope="module")
def lib():
    return mini_b200.load_library()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_numpy_mirror_matches_c_abi(lib, world):
    rng = np.random.default_rng(world)
    vs = np.concatenate([np.arange(64), rng.integers(0, 1 << 31, 2000), [(1 << 31) - 1, (1 << 32) - 1]])
    own, row = PT.owner(vs, world), PT.row(vs, world)
    for v, o, r in zip(vs.tolist(), own.tolist(), row.tolist()):
        co, cr, back = C.c_int32(), C.c_int64(), C.c_int64()
        assert lib.b200_partition_locate(world, v, C.byref(co), C.byref(cr)) == 0
        assert (co.value, cr.value) == (o, r)
        assert lib.b200_partition_global_id(world, co.value, cr.value, C.byref(back)) == 0
        assert back.value == v


def test_c_abi_rejects_bad_arguments(lib):
    o, r = C.c_int32(), C.c_int64()
    assert lib.b200_partition_locate(3, 5, C.byref(o), C.byref(r)) != 0      # not a power of two
    assert lib.b200_partition_locate(16, 5, C.byref(o), C.byref(r)) != 0     # more ranks than boxes
    assert lib.b200_partition_locate(2, -1, C.byref(o), C.byref(r)) != 0
    assert lib.b200_partition_global_id(4, 4, 0, C.byref(r)) != 0


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_bijection_equal_rows_per_rank(world):
    n = 1 << 16
    v = np.arange(n)
    own = PT.owner(v, world).astype(np.int64)
    assert np.array_equal(np.bincount(own, minlength=world), np.full(world, n // world))
    seen = np.zeros(n, bool)
    for r in range(world):
        ids = PT.global_ids(r, world, n // world)
        assert np.array_equal(PT.owner(ids, world), np.full(n // world, r, np.uint64))
        assert np.array_equal(PT.row(ids, world), np.arange(n // world, dtype=np.uint64))
        seen[ids] = True
    assert seen.all()
    assert np.array_equal(np.sort(PT.bit(v, world, n // world).astype(np.int64)), v)


@pytest.mark.parametrize("scale,world,tol", [(16, 8, 1.10), (22, 2, 1.01), (22, 8, 1.02)])
def test_arc_balance_on_rmat_degree_law(scale, world, tol):
    """Arc-weighted share of every rank under the RMAT marginal (each id bit is 0 w.p. a+b = 0.76):
    plain cyclic ownership gives rank 0 of 8 a 3.5x share, the swizzle stays within `tol` of 1/P."""
    w = np.ones(1)
    for _ in range(scale):
        w = np.concatenate([w * 0.76, w * 0.24])
    v = np.arange(1 << scale)
    share = np.bincount(PT.owner(v, world).astype(np.int64), weights=w, minlength=world) * world
    cyclic = np.bincount(v & (world - 1), weights=w, minlength=world) * world
    assert share.max() < tol and share.min() > 2 - tol
    assert cyclic.max() > 1.5
