"""Tiny driver for ncu: builds the RMAT graph and runs the chosen primitive a few times.
   python profiles/prof_bfs.py [--scale 22] [--runs 2] [--prim bfs|bfs_beamer|sssp|pr]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mini_b200 as mb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--prim", default="bfs")
ap.add_argument("--loop", default="graph", choices=["graph", "host"])
ap.add_argument("--timing", action="store_true", help="per-level CUDA-event timings (host loop)")
a = ap.parse_args()
ctx = mb.Context(0)
ctx.set_level_loop(mb.LOOP_HOST if a.loop == "host" else mb.LOOP_GRAPH)
g = ctx.prepare_graph(ctx.rmat_graph(a.scale, 16, 1, weighted=a.prim == "sssp"))
for _ in range(a.runs):
    if a.prim == "bfs":
        _, st = ctx.bfs(g, 0, mb.BFS_PUSH, timing=a.timing)
    elif a.prim == "bfs_beamer":
        _, st = ctx.bfs(g, 0, mb.BFS_BEAMER, 15.0, 18.0, timing=a.timing)
    elif a.prim == "sssp":
        _, st = ctx.sssp(g, 0)
    else:
        *_, st = ctx.pr(g, 3, True)
print(a.prim, "levels", st.num_levels, "device_ms", st.device_ms, "launches", st.launches)
for l in st.levels:
    print(l)
