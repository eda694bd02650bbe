"""CPU model of the near-far SSSP order (mini_b200/csrc/near_far.cuh) in NumPy: the same rule for which
improved vertices join the next frontier, the same bucket choice (smallest pending distance), the same width policy
(nf_open_bucket).  It pins the ALGORITHM on the CPU -- any bucket width reaches the oracle's distances, and the default
width keeps the relaxed arcs within 1.2x of the reached arcs (SURVEY 8f-4's bar) -- while the kernels themselves are
checked on the GPU (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import oracle

FLT_MAX = np.finfo(np.float32).max


def near_far(g: oracle.CSR, src: int, delta0: float):
    """Synchronous sweeps (the GPU's in-iteration atomicMin only makes improvements visible earlier).
    Returns (dist, iterations, buckets, relaxed_arcs)."""
    off, idx, w = g.offsets, g.indices, g.weights
    n, m = g.n, g.m
    dist = np.full(n, FLT_MAX, np.float32)
    dist[src] = 0
    frontier = np.array([src], np.int64)
    delta, cutoff = np.float32(delta0), np.float32(delta0)
    iters = buckets = 0
    relaxed = bucket_arcs = total_arcs = 0
    peaked = False
    while True:
        while frontier.size:                                     # near iterations of the open bucket
            s, cnt = off[frontier], off[frontier + 1] - off[frontier]
            tot = int(cnt.sum())
            iters += 1
            relaxed += tot
            bucket_arcs += tot
            if tot == 0:
                break
            seg = np.repeat(np.arange(frontier.size), cnt)
            pos = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt) + s[seg]
            with np.errstate(over="ignore"):
                nd = (dist[frontier][seg] + w[pos]).astype(np.float32)
            old = dist.copy()
            np.minimum.at(dist, idx[pos], nd)
            improved = np.nonzero(dist < old)[0]
            frontier = improved[dist[improved] < cutoff]         # SsspRelaxQ::finish: the others stay pending
        pending = dist[(dist >= cutoff) & (dist < FLT_MAX)]      # sssp_pending_min_kernel
        if pending.size == 0:
            return dist, iters, buckets, relaxed
        total_arcs += bucket_arcs                                # nf_open_bucket
        if total_arcs * 5 > m * 3:
            delta = np.float32(1e30)
        elif bucket_arcs * 32 < m:
            delta = np.float32(min(float(delta) * (4.0 if peaked else 2.0), 1e30))
        elif bucket_arcs * 4 > m:
            peaked, delta = True, np.float32(delta0)
        lo = pending.min()
        with np.errstate(over="ignore"):
            hi = np.float32(lo + delta)
        if not hi > lo:
            hi = np.nextafter(lo, np.float32(np.inf), dtype=np.float32)
        hi = min(hi, FLT_MAX)
        cutoff, bucket_arcs = hi, 0
        buckets += 1
        frontier = np.nonzero((dist >= lo) & (dist < hi))[0]     # sssp_take_kernel


def auto_delta(g: oracle.CSR) -> float:
    step = max(1, g.m // 65536)
    return 3.5 * float(g.weights[::step].astype(np.float64).mean()) * g.n / g.m


@pytest.mark.parametrize("delta", ["auto", float("inf"), 1e-3, 1.0, 7.5, 64.0, 1e9])
def test_any_bucket_width_reaches_the_oracle_distances(delta):
    g = oracle.rmat_csr(12, 16, 1, weighted=True)
    d0 = auto_delta(g) if delta == "auto" else delta
    for src in (0, 5):
        dist, iters, buckets, _ = near_far(g, src, d0)
        assert dist.tobytes() == oracle.sssp_dist(g, src).tobytes()
        assert iters >= 2
    rng = np.random.default_rng(3)
    gf = oracle.CSR(g.n, g.offsets, g.indices, (rng.random(g.m, dtype=np.float32) * 3.0 + 0.01).astype(np.float32))
    dist, *_ = near_far(gf, 0, auto_delta(gf) if delta == "auto" else delta)
    assert dist.tobytes() == oracle.sssp_dist_f32(gf, 0).tobytes()


@pytest.mark.parametrize("scale", [14, 16])
def test_default_width_relaxes_about_each_reached_arc_once(scale):
    g = oracle.rmat_csr(scale, 16, 1, weighted=True)
    want = oracle.sssp_dist(g, 0)
    reached = int(np.diff(g.offsets)[want < FLT_MAX].sum())
    dist, iters, buckets, relaxed = near_far(g, 0, auto_delta(g))
    assert dist.tobytes() == want.tobytes()
    assert reached <= relaxed <= 1.2 * reached, (relaxed, reached)
    dist_bf, iters_bf, buckets_bf, relaxed_bf = near_far(g, 0, float("inf"))       # the reference's order
    assert dist_bf.tobytes() == want.tobytes() and buckets_bf == 0
    assert relaxed_bf > 1.5 * reached and relaxed < relaxed_bf
    assert iters <= 3 * iters_bf                                                   # the price: a few more iterations
