"""Kernel-entry timeline (B200_LOOP_TRACE) of the single-GPU graph-driven loops: push BFS, direction-optimising BFS and
SSSP on the bench's RMAT graph.  The traced run is the last one of each (warm graph, warm caches).
    python profiles/trace_single.py [--scale 22] 2> trace.err;  python profiles/format_trace.py trace.err single"""
import argparse
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
import mini_b200 as mb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=22)
a = ap.parse_args()
os.environ.pop("B200_LOOP_TRACE", None)
ctx = mb.Context(0)
g = ctx.rmat_graph(a.scale, 16, 1, weighted=True)
labels = torch.empty(g.n, dtype=torch.int32, device="cuda")
dist = torch.empty(g.n, dtype=torch.float32, device="cuda")
runs = [("push BFS", lambda: ctx.bfs(g, 0, mb.BFS_PUSH, labels=labels)),
        ("direction-optimising BFS", lambda: ctx.bfs(g, 0, mb.BFS_BEAMER, alpha=15.0, beta=18.0, labels=labels)),
        ("SSSP (near-far order)", lambda: ctx.sssp(g, 0, dist=dist))]
for name, fn in runs:
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    sys.stderr.write(f"# {name}, RMAT scale-{a.scale}\n")
    sys.stderr.flush()
    os.environ["B200_LOOP_TRACE"] = "1"
    fn()
    os.environ.pop("B200_LOOP_TRACE")
