"""A/B helper: per-level advance times of a push BFS (and SSSP) for the library named by
B200_FRONTIER_LIB (tuning builds, see mini_b200/build.py --define/--out).
   python profiles/ab_levels.py [--scale 22] [--reps 10] [--advance quad|lbs] [--sssp]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mini_b200 as mb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--advance", default="quad")
ap.add_argument("--sssp", action="store_true")
a = ap.parse_args()
ctx = mb.Context(0)
ctx.set_advance_impl(mb.ADVANCE_LBS if a.advance == "lbs" else mb.ADVANCE_QUAD)
g = ctx.rmat_graph(a.scale, 16, 1, weighted=a.sssp)
tag = os.path.basename(os.environ.get("B200_FRONTIER_LIB", "default")) + "/" + a.advance
for mode, name in ((mb.BFS_PUSH, "push"), (mb.BFS_BEAMER, "beamer")):
    ms = None
    tot = 0.0
    for r in range(a.reps + 2):
        _, st = ctx.bfs(g, 0, mode, 15.0, 18.0, timing=True)
        if r < 2:
            continue
        if ms is None:
            ms = [0.0] * len(st.levels)
        for i, l in enumerate(st.levels):
            ms[i] += l["advance_ms"] / a.reps
        tot += st.device_ms / a.reps
    print(tag, name, "total_ms(with timing events) %.4f" % tot, "levels:",
          " ".join("%s%.4f" % ("v" if l["direction"] else "^", m) for l, m in zip(st.levels, ms)))
    t = 0.0
    for r in range(a.reps):
        _, st = ctx.bfs(g, 0, mode, 15.0, 18.0)
        t += st.device_ms / a.reps
    print(tag, name, "device_ms %.4f" % t)
if a.sssp:
    t = 0.0
    for r in range(3):
        _, st = ctx.sssp(g, 0)
        t += st.device_ms / 3
    print(tag, "sssp device_ms %.3f iters %d arcs %d" % (t, st.num_levels, st.total_arcs))
