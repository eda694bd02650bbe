"""Partitioned (multi-GPU) BFS.
 * On ONE GPU: P virtual ranks run as threads over ThreadComm, exercising the partition-aware kernels
   (swizzled-cyclic owner mapping, rank-major bitmaps, in-kernel routing, absorb, partitioned pull) bit-exactly
   against the CPU oracle.
 * With >= 2 GPUs (gpurun --gpus N): the same driver over NCCL (tests/run_dist_bfs.py under torchrun)."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import oracle
from mini_b200 import partition as PT

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _virtual_ranks(scale, ef, seed, world, src, mode):
    import mini_b200 as mb
    from mini_b200 import dist as D
    shared = D.ThreadComm.Shared(world)
    out, errs = [None] * world, []

    def work(rank):
        try:
            ctx = mb.Context(0)
            g = D.build_rank_graph(ctx, scale, ef, seed, rank, world)
            n = 1 << scale
            rk = D.GpuRank(ctx, rank, world, n, g)
            bfs = D.DistBFS(rk, D.ThreadComm(shared, rank), n, (2 * ef) << scale, mode=mode)
            levels = bfs.run(src)
            torch.cuda.synchronize()
            out[rank] = (rk.labels.cpu().numpy(), g.offsets_host(), g.col_indices.cpu().numpy(), levels, bfs.levels)
            ctx.close()
        except BaseException as e:   # noqa: BLE001
            errs.append(e)
            shared.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(600) for t in ts]
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("mode", ["push", "beamer"])
def test_partitioned_bfs_virtual_ranks(world, mode):
    scale, ef, seed, src = 14, 16, 1, 0
    o = oracle.rmat_csr(scale, ef, seed)
    ref = oracle.bfs(o, src)
    res = _virtual_ranks(scale, ef, seed, world, src, mode)
    n = 1 << scale
    labels = np.empty(n, np.int32)
    for r, (lab, off, idx, levels, stats) in enumerate(res):
        rows = PT.global_ids(r, world, n // world)   # global id of every local row (swizzled-cyclic partition)
        labels[rows] = lab
        # the rank's CSR is exactly its rows of the global CSR
        assert np.array_equal(np.diff(off), (o.offsets[rows + 1] - o.offsets[rows]))
        k = min(len(rows), 50)
        for i in range(k):
            assert np.array_equal(idx[off[i]:off[i + 1]], o.indices[o.offsets[rows[i]]:o.offsets[rows[i] + 1]])
    assert np.array_equal(labels, ref)
    stats = res[0][4]
    assert sum(l["discovered"] for l in stats) + 1 == int((ref >= 0).sum())
    if mode == "beamer":
        assert any(l["direction"] == "pull" for l in stats)
    if world > 1 and mode == "push":
        # each remote vertex is sent at most once per sender
        assert sum(l["sent"] for l in stats) <= (world - 1) * n


@pytest.mark.parametrize("src", [1, 77, 12345])
def test_partitioned_bfs_other_sources(src):
    scale, ef, seed, world = 14, 8, 3, 4
    ref = oracle.bfs(oracle.rmat_csr(scale, ef, seed), src)
    res = _virtual_ranks(scale, ef, seed, world, src, "beamer")
    labels = np.empty(1 << scale, np.int32)
    for r in range(world):
        labels[PT.global_ids(r, world, (1 << scale) // world)] = res[r][0]
    assert np.array_equal(labels, ref)


SMALL_VARIANTS = {   # B200_P2P_SMALL / _SMALL_ARCS / _SMALL_VERTS / _HUB_MIN (p2p_bfs.cu): which levels the persistent small-level kernel runs
    "small-default": ("1", None, None, "64"),    # scale-14 graphs: every push level is "small", except level 0 from a source of >= 64 arcs
    "small-tiny": ("1", "3000", "64", None),     # only the first / last levels: exercises small -> big push -> pull -> small
    "small-off": ("0", None, None, None),        # every push level through scan + claim-only advance + bitmap exchange
}


def _virtual_ranks_p2p(scale, ef, seed, world, src, mode, repeat=2, loop="graph", small="small-default"):
    """P ranks as threads of this process, each with its OWN stream on GPU 0 (the cross-rank flag
    barriers spin inside kernels, so the ranks' kernels must be able to run concurrently); heaps are
    wired by address (b200_p2p_bfs_connect with peer_bases)."""
    import mini_b200 as mb
    on, arcs, verts, hub = SMALL_VARIANTS[small]
    for k, v in (("B200_P2P_SMALL", on), ("B200_P2P_SMALL_ARCS", arcs), ("B200_P2P_SMALL_VERTS", verts), ("B200_P2P_HUB_MIN", hub)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    # the small-level kernel synchronises its own grid: the grids of all ranks sharing this GPU must be resident together
    os.environ["B200_P2P_SMALL_GRID"] = str(max(1, 128 // world))
    os.environ["B200_P2P_PULL_GRID"] = str(max(1, 256 // world))
    from mini_b200 import dist as D
    from mini_b200.p2p import P2PBfs
    bar = threading.Barrier(world)
    bases, out, errs = [None] * world, [None] * world, []
    n = 1 << scale

    def work(rank):
        try:
            ctx = mb.Context(0, stream=None)
            ctx.set_level_loop(mb.LOOP_GRAPH if loop == "graph" else mb.LOOP_HOST)
            g = D.build_rank_graph(ctx, scale, ef, seed, rank, world)
            if world in (2, 8):
                g = ctx.prepare_graph(g)   # caller-owned "no in-arc" bitmap; the other worlds let the engine derive it
            bfs = P2PBfs(ctx, rank, world, n, (2 * ef) << scale, g)
            bases[rank] = bfs.heap_base
            bar.wait()
            if world > 1:
                bfs.connect_local(bases)
            bfs.prepare(mode)        # graph-driven loop: build the traversal graph before any rank spins in a barrier
            bar.wait()
            for _ in range(repeat):
                levels = bfs.run(src, mode)
                assert bfs.level_loop == loop, "the requested level loop did not run"
            ctx.sync()
            out[rank] = (bfs.labels.cpu().numpy(), levels, list(bfs.levels))
            bar.wait()
            bfs.close()
            ctx.close()
        except BaseException as e:   # noqa: BLE001
            errs.append(e)
            bar.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(600) for t in ts]
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("loop,small", [("graph", "small-default"), ("graph", "small-tiny"), ("graph", "small-off"),
                                        ("host", "small-default")])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("mode", ["push", "beamer"])
def test_p2p_bfs_virtual_ranks(world, mode, loop, small):
    scale, ef, seed, src = 14, 16, 1, 0
    ref = oracle.bfs(oracle.rmat_csr(scale, ef, seed), src)
    res = _virtual_ranks_p2p(scale, ef, seed, world, src, mode, loop=loop, small=small)
    labels = np.empty(1 << scale, np.int32)
    for r in range(world):
        labels[PT.global_ids(r, world, (1 << scale) // world)] = res[r][0]
    assert np.array_equal(labels, ref)
    stats = res[0][2]
    assert all(res[r][2] == stats or [dict(l, level_ms=0) for l in res[r][2]] == [dict(l, level_ms=0) for l in stats]
               for r in range(world))                    # every rank saw the same global level summary
    assert sum(l["discovered"] for l in stats) + 1 == int((ref >= 0).sum())
    if mode == "beamer":
        assert any(l["direction"] == "pull" for l in stats)
    if world > 1 and mode == "push" and (loop == "host" or small != "small-off"):
        assert 0 < sum(l["sent"] for l in stats) <= (world - 1) * (1 << scale)
    if loop == "graph":
        kinds = {l["exchange"] for l in stats if l["direction"] == "push"}
        # (small-default sets B200_P2P_HUB_MIN=64: with more than one rank, level 0 from the hub goes through the bitmap exchange)
        want = {"small-default": {"ids"} if world == 1 else {"ids", "bitmap"}, "small-tiny": {"ids", "bitmap"}, "small-off": {"bitmap"}}[small]
        assert kinds == want, kinds


@pytest.mark.parametrize("small", ["small-default", "small-tiny"])
@pytest.mark.parametrize("src", [1, 77, 12345])
def test_p2p_bfs_other_sources(src, small):
    scale, ef, seed, world = 14, 8, 3, 4
    ref = oracle.bfs(oracle.rmat_csr(scale, ef, seed), src)
    res = _virtual_ranks_p2p(scale, ef, seed, world, src, "beamer", small=small)
    labels = np.empty(1 << scale, np.int32)
    for r in range(world):
        labels[PT.global_ids(r, world, (1 << scale) // world)] = res[r][0]
    assert np.array_equal(labels, ref)


def test_p2p_ranks_if_multiple_gpus():
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if ngpu < 4 else (4 if ngpu < 8 else 8)
    for mode in ("beamer", "push"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
               "127.0.0.1", "--master-port", "29621", os.path.join(ROOT, "tests", "run_p2p_bfs.py"), "--scale", "18",
               "--mode", mode]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "P2P_BFS_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_nccl_ranks_if_multiple_gpus():
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "run_dist_bfs.py"), "--scale", "16", "--backend", "nccl"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_BFS_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
