"""Driver for ncu / timelines: the peer-memory BFS with ONE rank (world = 1), so that its persistent kernels
(p2p_small_levels_kernel, p2p_pull_levels_kernel) can be captured on a single GPU.
   python profiles/prof_p2p_single.py [--scale 25] [--mode beamer] [--runs 2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mini_b200 as mb  # noqa: E402
from mini_b200 import dist as D  # noqa: E402
from mini_b200.p2p import P2PBfs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=25)
ap.add_argument("--mode", default="beamer")
ap.add_argument("--runs", type=int, default=2)
a = ap.parse_args()
ctx = mb.Context(0)
g = ctx.prepare_graph(D.build_rank_graph(ctx, a.scale, 16, 1, 0, 1))
rk = P2PBfs(ctx, 0, 1, 1 << a.scale, 32 << a.scale, g)
rk.prepare(a.mode)
rk.set_trace(True)
for _ in range(a.runs):
    rk.run(0, a.mode)
    print("ms", rk.device_ms, [(round(t, 1), k) for t, k in rk.last_trace()])
    print([dict(d=l["direction"][:4], x=l["exchange"], F=l["frontier"], arcs=l["arcs"], found=l["discovered"]) for l in rk.levels])
# the single-GPU engine on the same graph: per-level times of the host-driven loop (CUDA events)
for _ in range(3):
    _, st = ctx.bfs(g, 0, mb.BFS_BEAMER if a.mode == "beamer" else mb.BFS_PUSH, 15.0, 18.0, timing=True)
print("engine", st.device_ms, [("pull" if l["direction"] else "push", l["frontier_len"], l["arcs"], l["discovered"],
                               round(l["advance_ms"], 4), round(l["level_ms"], 4)) for l in st.levels])
rk.close()
ctx.close()
