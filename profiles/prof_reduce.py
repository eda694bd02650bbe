"""Tiny driver for ncu: the full-frontier neighbourhood reduce of BASELINE.json configs[4] (RMAT scale-24, fp32 plus).
   python profiles/prof_reduce.py [--scale 24] [--runs 3] [--hot 40960|0]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mini_b200 as mb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--runs", type=int, default=3)
ap.add_argument("--hot", type=int, default=40960)
a = ap.parse_args()
ctx = mb.Context(0)
g = ctx.rmat_graph(a.scale, 16, 1)
if a.hot:
    ctx.prepare_hot_columns(g, a.hot)
vals = torch.rand(g.n, device="cuda")
frontier = torch.arange(g.n, dtype=torch.int32, device="cuda")
red = torch.empty(g.n, dtype=torch.float32, device="cuda")
for _ in range(a.runs):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    arcs = ctx.neighborhood_reduce(g, frontier, vals, red, 0.0)
    e1.record()
    torch.cuda.synchronize()
    print("reduce", "hot" if a.hot else "plain", "arcs", arcs, "ms (scan + reduce + read-back)", e0.elapsed_time(e1))
hot_frac = float((g.hot_indices < 0).float().mean().item()) if a.hot else 0.0
print("fraction of arcs that point at a hot column:", hot_frac)
ctx.close()
