"""bench.py leg for N > 1 GPUs: direction-optimising BFS on RMAT scale-26 ef16 (BASELINE.json
configs[3]), cyclic 1D vertex partition, one rank per GPU, NCCL frontier exchange.  Strong scaling:
the graph is fixed, each rank holds 1/N of the rows."""
from __future__ import annotations

import json
import os
import time


def run(args) -> int:
    import torch
    import torch.distributed as dist
    import mini_b200 as mb
    from mini_b200 import dist as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl.%h.%p.log")   # keep NCCL's banner off stdout (one JSON line)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    scale = args.scale or 26
    ef = 16
    n, m = 1 << scale, (2 * ef) << scale
    ctx = mb.Context(dev)
    t0 = time.time()
    g = D.build_rank_graph(ctx, scale, ef, 1, rank, world)
    build_s = time.time() - t0
    rk = D.GpuRank(ctx, rank, world, n, g)
    comm = D.TorchComm(ctx.torch_device)
    mode = "beamer" if args.mode in ("push", "beamer") and args.mg_mode == "beamer" else "push"
    bfs = D.DistBFS(rk, comm, n, m, mode=mode)

    sampler = None
    if rank == 0:
        from bench import ClockSampler
        sampler = ClockSampler(dev)
        sampler.start()

    def timed(steps, fn):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = ctx_launches()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=ctx.torch_device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx_launches() - launches0

    import ctypes as C
    from mini_b200 import lib as L

    class _WS(C.Structure):   # prefix of b200_workspace up to `launches`
        _fields_ = [("stream", C.c_void_p), ("device", C.c_int32), ("num_sms", C.c_int32), ("d_status", C.c_void_p),
                    ("status_tiles", C.c_int64), ("d_tile_counter", C.c_void_p), ("epoch", C.c_uint), ("reserved", C.c_uint),
                    ("d_counters", C.c_void_p), ("h_counters", C.c_void_p), ("d_scanned", C.c_void_p),
                    ("scanned_capacity", C.c_int64), ("launches", C.c_int64)]

    ws = C.cast(L.load_library().b200_ctx_workspace(ctx._h), C.POINTER(_WS))

    def ctx_launches():
        return int(ws.contents.launches)

    for _ in range(args.warmup):
        bfs.run(0)
    ms_total, launches = timed(args.steps, lambda: bfs.run(0))
    levels = list(bfs.levels)
    reached = comm.all_reduce_sum([rk.reached_degree_sum(), rk.reached_count()])
    reached_arcs, reached_vertices = reached
    value = reached_arcs * args.steps / (ms_total * 1e-3) / 1e9

    # end to end: every step uploads the rank's initial labels from pinned host memory, runs, downloads labels
    h_init = torch.full((rk.n_local,), -1, dtype=torch.int32).pin_memory()
    h_out = torch.empty(rk.n_local, dtype=torch.int32).pin_memory()

    def e2e_step():
        rk.labels.copy_(h_init, non_blocking=True)
        bfs.run(0)
        h_out.copy_(rk.labels, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=ctx.torch_device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    props_ok, level_hist = D.verify_bfs_properties(rk, comm, 0)
    all_launches = comm.all_reduce_sum([launches])[0]
    sent = sum(l["sent"] for l in levels)
    if rank == 0:
        clocks = sampler.stop()
        line = {
            "metric": "bfs_gteps_rmat", "value": value, "unit": "GTEPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"direction-optimising BFS from vertex 0, RMAT scale-{scale} ef16 symmetrised "
                                   f"(n={n}, m={m}), cyclic 1D vertex partition over {world} GPUs, NCCL alltoallv "
                                   "(push) / allgather of bitmap slices (pull)",
                       "mode": mode, "teps_numerator": "sum of deg(v) over reached v", "reached_arcs": reached_arcs,
                       "reached_vertices": reached_vertices, "parallelism": f"1d-cyclic x{world}",
                       "l2": "inputs larger than L2 (per-rank col_indices %d MiB)" % (g.m * 4 >> 20),
                       "graph_build_s": round(build_s, 2),
                       "levels": [dict(d=l["direction"], F=l["frontier"], arcs=l["arcs"], found=l["discovered"], sent=l["sent"])
                                  for l in levels]},
            "clocks": clocks,
            "e2e": {"value": reached_arcs * args.steps / e2e_s / 1e9, "unit": "GTEPS",
                    "h2d_bytes_per_step": rk.n_local * 4 * world, "d2h_bytes_per_step": rk.n_local * 4 * world,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": all_launches,
            "roofline": {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "per-kernel roofline is reported by the N=1 line; exchange volume: "
                                 f"{sent * 4} bytes of vertex ids per BFS over NVLink"},
            "cpu_baseline": None,
            "parity": {"bfs_properties_hold_at_full_scale": props_ok, "vertices_per_level": level_hist},
        }
        print(json.dumps(line))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    return 0 if props_ok else 1
