"""Per-iteration table of b200_sssp_run on the bench's SSSP workload (host-driven loop with CUDA events), near-far order
against the reference's Bellman-Ford order, and the graph-loop wall time of both.
    python profiles/sssp_levels.py [--scale 22] [--delta X ...]"""
import argparse
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
import mini_b200 as mb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--delta", type=float, nargs="*", default=[0.0, float("inf")])
a = ap.parse_args()
ctx = mb.Context(0)
g = ctx.rmat_graph(a.scale, 16, 1, weighted=True)
dist = torch.empty(g.n, dtype=torch.float32, device="cuda")
off = g.row_offsets.to(torch.int64) & 0xFFFFFFFF
for delta in a.delta:
    ctx.set_sssp_delta(delta)
    for _ in range(3):
        ctx.sssp(g, 0, dist=dist)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _, st = ctx.sssp(g, 0, dist=dist)
    e1.record()
    torch.cuda.synchronize()
    reached = int((off[1:] - off[:-1])[dist < 3.0e38].sum().item())
    _, tt = ctx.sssp(g, 0, dist=dist, timing=True)
    print(json.dumps({"delta": delta, "graph_loop_ms": e0.elapsed_time(e1) / 10, "iterations": st.num_levels,
                      "relaxed_arcs": st.total_arcs, "reached_arcs": reached, "ratio": st.total_arcs / reached,
                      "host_loop_levels": [[l["frontier_len"], l["arcs"], l["discovered"], round(l["advance_ms"], 4),
                                            round(l["level_ms"], 4)] for l in tt.levels]}))
