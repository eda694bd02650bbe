"""torchrun entry (GPUs only): peer-memory BFS (mini_b200.p2p, b200_p2p_bfs_*) with one rank per GPU,
IPC handles traded over NCCL, labels checked against the CPU oracle on rank 0."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=16)
ap.add_argument("--mode", default="beamer")
ap.add_argument("--src", type=int, default=0)
ap.add_argument("--repeat", type=int, default=3)
a = ap.parse_args()

import torch
import torch.distributed as dist

import mini_b200 as mb
import oracle
from mini_b200 import partition as PT
from mini_b200 import dist as D
from mini_b200.p2p import P2PBfs

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
n, ef = 1 << a.scale, 16
ctx = mb.Context(dev)
g = D.build_rank_graph(ctx, a.scale, ef, 1, rank, world)
bfs = P2PBfs(ctx, rank, world, n, (2 * ef) << a.scale, g)
bfs.connect_torch_distributed()
bfs.prepare(a.mode)
dist.barrier()
for _ in range(a.repeat):               # repeated runs reuse the heap: epochs / double buffers must stay consistent
    levels = bfs.run(a.src, a.mode)
gl = [torch.empty_like(bfs.labels) for _ in range(world)]
dist.all_gather(gl, bfs.labels)
if rank == 0:
    full = np.empty(n, np.int32)
    for r in range(world):
        full[PT.global_ids(r, world, n // world)] = gl[r].cpu().numpy()
    o = oracle.rmat_csr(a.scale, ef, 1)
    ref = oracle.bfs(o, a.src)
    assert np.array_equal(full, ref), "labels differ from the oracle"
    dirs = "".join("P" if l["direction"] == "pull" else "p" for l in bfs.levels)
    print(f"P2P_BFS_OK world={world} levels={levels} dirs={dirs} sent={sum(l['sent'] for l in bfs.levels)} ms={bfs.device_ms:.3f}")
dist.barrier()
bfs.close()
ctx.close()
dist.destroy_process_group()
