"""Workload for compute-sanitizer (profiles/sanitize.sh): every product kernel family once, small enough for the
tools' 10-100x slow-down, results checked against the CPU oracle so a "clean" log is a log of a CORRECT run.
  BFS push / direction-optimising (graph-driven and host-driven level loops), SSSP (+ deterministic preds),
  PR-style neighbourhood reduce (plain and hot-column), operator-level compaction / uniquify, one-rank peer-memory BFS
  (small-level kernel, pull-levels kernel, bitmap absorb)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mini_b200 as mb
import oracle
from mini_b200 import dist as D
from mini_b200.p2p import P2PBfs

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 14
loops = sys.argv[2] if len(sys.argv) > 2 else "both"      # host | graph | both
with_p2p = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
ctx = mb.Context(0)
g = ctx.prepare_graph(ctx.rmat_graph(scale, 16, 1, weighted=True))
o = oracle.CSR(g.n, g.offsets_host(), g.col_indices.cpu().numpy(), g.col_values.cpu().numpy())
ref = oracle.bfs(o, 0)
for loop in [l for l, name in ((mb.LOOP_GRAPH, "graph"), (mb.LOOP_HOST, "host")) if loops in (name, "both")]:
    ctx.set_level_loop(loop)
    for mode in (mb.BFS_PUSH, mb.BFS_BEAMER, mb.BFS_REF_ALPHA):
        lab, _ = ctx.bfs(g, 0, mode, 15.0 if mode != mb.BFS_REF_ALPHA else 2.0, 18.0)
        assert np.array_equal(lab.cpu().numpy(), ref)
    preds = torch.empty(g.n, dtype=torch.int32, device="cuda")
    for delta in (0.0, 1.0, float("inf")):   # near-far order: automatic bucket width, one bucket per distance, the reference's order
        ctx.set_sssp_delta(delta)
        dist, _ = ctx.sssp(g, 0, preds=preds)
        assert dist.cpu().numpy().tobytes() == oracle.sssp_dist(o, 0).tobytes()
    ctx.set_sssp_delta(0.0)
cur, red, lens, _ = ctx.pr(g, 3, False)
ocur, _, olens = oracle.pr(o, 3, False)
assert list(lens) == olens.tolist() and np.allclose(cur.cpu().numpy(), ocur, rtol=1e-4, atol=1e-6)
ctx.prepare_hot_columns(g, hot_count=4096)
cur, red, lens, _ = ctx.pr(g, 3, False)
assert list(lens) == olens.tolist() and np.allclose(cur.cpu().numpy(), ocur, rtol=1e-4, atol=1e-6)
# one-rank peer-memory BFS: the persistent kernels with their grid barriers (no peer, so no flag waits)
for small in (("1", "0") if with_p2p else ()):
    os.environ["B200_P2P_SMALL"] = small
    gp = ctx.prepare_graph(D.build_rank_graph(ctx, scale, 16, 1, 0, 1))
    rk = P2PBfs(ctx, 0, 1, g.n, g.m, gp)
    for mode in ("beamer", "push"):
        rk.prepare(mode)
        rk.run(0, mode)
        assert np.array_equal(rk.labels.cpu().numpy(), ref), (small, mode)
    rk.close()
ctx.close()
print("SANITIZE_RUN_OK scale", scale, "loops", loops, "p2p", with_p2p)
