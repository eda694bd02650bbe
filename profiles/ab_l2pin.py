"""A/B of a persisting-L2 access-policy window over the visited bitmap (B200_L2_PIN_VISITED=1, host-driven loop) for
push BFS: the bitmap is 512 KiB at scale 22 (L1-resident by design) and 8 MiB at scale 26 (L2 only).
    python profiles/ab_l2pin.py [--scale 26]"""
import argparse
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
import mini_b200 as mb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=26)
a = ap.parse_args()
for pin in ("0", "1", "0", "1"):
    os.environ["B200_L2_PIN_VISITED"] = pin
    ctx = mb.Context(0)
    ctx.set_level_loop(mb.LOOP_HOST)
    g = ctx.rmat_graph(a.scale, 16, 1)
    labels = torch.empty(g.n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.bfs(g, 0, mb.BFS_PUSH, labels=labels)
    ms, lv = 0.0, None
    for _ in range(5):
        _, st = ctx.bfs(g, 0, mb.BFS_PUSH, labels=labels, timing=True)
        ms += st.device_ms / 5
        lv = [round(l["advance_ms"], 4) for l in st.levels]
    print(f"scale {a.scale} pin={pin}: device_ms {ms:.4f} advance_ms per level {lv}", flush=True)
    ctx.l2_pin(None)
    ctx.close()
    del g, labels
