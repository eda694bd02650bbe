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
a = ap.parse_args()
ctx = mb.Context(0)
g = ctx.rmat_graph(a.scale, 16, 1, weighted=a.prim == "sssp")
for _ in range(a.runs):
    if a.prim == "bfs":
        _, st = ctx.bfs(g, 0, mb.BFS_PUSH)
    elif a.prim == "bfs_beamer":
        _, st = ctx.bfs(g, 0, mb.BFS_BEAMER, 15.0, 18.0)
    elif a.prim == "sssp":
        _, st = ctx.sssp(g, 0)
    else:
        *_, st = ctx.pr(g, 3, True)
print(a.prim, "levels", st.num_levels, "device_ms", st.device_ms, "launches", st.launches)
for l in st.levels:
    print(l)
