"""torchrun entry: kernel-entry timeline (B200_LOOP_TRACE) and per-BFS times of the peer-memory multi-GPU BFS.

    B200_LOOP_TRACE=all python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29631 profiles/trace_p2p.py --scale 26 [--modes beamer,push] [--reps 10]

Every rank prints its own timeline to stderr (the library does, one line per traced run); rank 0 prints one JSON
line per mode with the max-over-ranks time per BFS and the level summary.  Trace ids (p2p_bfs.cu): 20 small-level
kernel entry, 21 level start (phase A), 22 sends out / wait for the peers' flags, 23 absorb, 24 stats posted /
wait, 25 decision; 2 quad scan, 3 quad advance, 5 bitmap absorb entry, 13 after its flag barrier, 7 gather + OR,
8 pull, 9 stats + decide entry, 12 after its barrier, 10 hand-over gather, 11 hand-over bitmap -> list."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=26)
ap.add_argument("--modes", default="beamer,push")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()

import torch
import torch.distributed as dist

import mini_b200 as mb
from mini_b200 import dist as D
from mini_b200.p2p import P2PBfs

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl.%h.%p.log")
dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
trace = os.environ.pop("B200_LOOP_TRACE", None)      # only the last run of every mode is traced
n, ef = 1 << a.scale, 16
ctx = mb.Context(dev)
g = ctx.prepare_graph(D.build_rank_graph(ctx, a.scale, ef, 1, rank, world))
bfs = P2PBfs(ctx, rank, world, n, (2 * ef) << a.scale, g)
bfs.connect_torch_distributed()
for mode in a.modes.split(","):
    bfs.prepare(mode)
    dist.barrier()
    for _ in range(3):
        bfs.run(0, mode)
    times = []
    for _ in range(a.reps):
        dist.barrier()
        torch.cuda.synchronize()
        bfs.run(0, mode)
        t = torch.tensor([bfs.device_ms], dtype=torch.float64, device=ctx.torch_device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    if trace:
        os.environ["B200_LOOP_TRACE"] = trace
        dist.barrier()
        bfs.run(0, mode)
        os.environ.pop("B200_LOOP_TRACE")
    if rank == 0:
        times.sort()
        print(json.dumps({"mode": mode, "world": world, "scale": a.scale, "ms_median": times[len(times) // 2], "ms_min": times[0],
                          "levels": [dict(d=l["direction"][:4], x=l["exchange"], F=l["frontier"], arcs=l["arcs"], found=l["discovered"],
                                          sent=l["sent"]) for l in bfs.levels]}), flush=True)
dist.barrier()
bfs.close()
ctx.close()
dist.destroy_process_group()
