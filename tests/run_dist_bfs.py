"""torchrun entry: partitioned BFS over torch.distributed, checked against the CPU oracle on rank 0.
   --backend nccl : GPUs, real kernels (GpuRank).   --backend gloo : CPU, NumPy stand-in (host logic only)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=12)
ap.add_argument("--backend", default="gloo")
ap.add_argument("--mode", default="beamer")
ap.add_argument("--src", type=int, default=0)
a = ap.parse_args()

import torch
import torch.distributed as dist

import oracle
from mini_b200 import partition as PT
from mini_b200 import dist as D

dist.init_process_group(a.backend)
rank, world = dist.get_rank(), dist.get_world_size()
n, ef = 1 << a.scale, 16
o = oracle.rmat_csr(a.scale, ef, 1)
if a.backend == "nccl":
    import mini_b200 as mb
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    ctx = mb.Context(dev)
    g = D.build_rank_graph(ctx, a.scale, ef, 1, rank, world)
    rk = D.GpuRank(ctx, rank, world, n, g)
    comm = D.TorchComm(ctx.torch_device)
else:
    from numpy_rank import NumpyRank
    rk = NumpyRank(o, rank, world)
    comm = D.TorchComm(torch.device("cpu"))
bfs = D.DistBFS(rk, comm, n, o.m, mode=a.mode)
levels = bfs.run(a.src)
lab = rk.labels if a.backend == "gloo" else rk.labels.cpu()
lab = torch.as_tensor(np.asarray(lab)).contiguous()
gathered = [torch.empty_like(lab) for _ in range(world)]
if a.backend == "nccl":
    gl = [t.cuda() for t in gathered]
    dist.all_gather(gl, lab.cuda())
    gathered = [t.cpu() for t in gl]
else:
    dist.all_gather(gathered, lab)
if rank == 0:
    full = np.empty(n, np.int32)
    for r in range(world):
        full[PT.global_ids(r, world, n // world)] = gathered[r].numpy()
    ref = oracle.bfs(o, a.src)
    assert np.array_equal(full, ref), "labels differ from the oracle"
    dirs = "".join("P" if l["direction"] == "pull" else "p" for l in bfs.levels)
    print(f"DIST_BFS_OK world={world} levels={levels} dirs={dirs} sent={sum(l['sent'] for l in bfs.levels)}")
dist.barrier()
dist.destroy_process_group()
