"""A/B of the WHILE-body unroll factor of the single-GPU graph loop (B200_LOOP_UNROLL, read when the graph is built):
wall time per traversal of push BFS, direction-optimising BFS and SSSP on the bench's RMAT graph.
    python profiles/ab_unroll.py [--scale 22]"""
import argparse
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
import mini_b200 as mb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--unroll", type=int, nargs="*", default=[1, 2, 3, 4, 1])
a = ap.parse_args()
for u in a.unroll:
    os.environ["B200_LOOP_UNROLL"] = str(u)
    ctx = mb.Context(0)
    g = ctx.rmat_graph(a.scale, 16, 1, weighted=True)
    labels = torch.empty(g.n, dtype=torch.int32, device="cuda")
    dist = torch.empty(g.n, dtype=torch.float32, device="cuda")
    out = []
    for name, fn in (("push", lambda: ctx.bfs(g, 0, mb.BFS_PUSH, labels=labels)),
                     ("do", lambda: ctx.bfs(g, 0, mb.BFS_BEAMER, alpha=15.0, beta=18.0, labels=labels)),
                     ("sssp", lambda: ctx.sssp(g, 0, dist=dist))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            _, st = fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{name} {e0.elapsed_time(e1) / 30:.4f} ms ({st.num_levels} levels)")
    print(f"unroll {u}: " + "  ".join(out), flush=True)
    ctx.close()
    del g, labels, dist
