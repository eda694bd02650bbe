"""bench.py leg for N > 1 GPUs (bench code, not product: it may run the CPU oracle as the checker): direction-optimising BFS on RMAT scale-26 ef16 (BASELINE.json
configs[3]), cyclic 1D vertex partition, one rank per GPU.  Strong scaling: the graph is fixed, each
rank holds 1/N of the rows.

Frontier exchange (`--exchange`):
  p2p  (default)  the exchange is fused into the kernels over NVLink peer memory (mini_b200.p2p,
                  csrc/p2p_bfs.cu): no collective call on the data path, one host sync per level;
  nccl            alltoallv / allgather / allreduce through torch.distributed (mini_b200.dist) -- the
                  baseline form of the same algorithm.
torch.distributed (NCCL) is used in both for rendezvous, IPC-handle exchange, barriers and the
max-over-ranks of the timings."""
from __future__ import annotations

import json
import os
import time


def _level_rows(levels):
    return [dict(d=l["direction"], F=l["frontier"], arcs=l["arcs"], found=l["discovered"], sent=l["sent"]) for l in levels]


def _mg_roofline(levels, level_ms, timeline, n, world, sent):
    """Per-GPU roofline of the heaviest pull level (B_pull = |U|*12 + a*(4 + 1/8) + |F_next|*4 + n/8,
    uint32 offsets; |U| = vertices still unlabelled when the level starts, a = in-arcs inspected, counted by the kernel)
    and the NVLink figure of its allgather (8d: (P-1)/P * n/8 bytes received per GPU over the gather's own time)."""
    from bench import _peaks
    peak, peak_src = _peaks()
    out = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
           "nvlink": {"vertex_ids_sent_bytes_per_bfs": sent * 4, "bitmap_bytes_gathered_per_pull_level_per_gpu": (world - 1) * (n // 8) // world,
                      "peak_gbs_per_direction": 900.0, "achieved_gbs": None, "frac": None}}
    pulls = [i for i, l in enumerate(levels) if l["direction"] == "pull"]
    if not pulls or len(level_ms) != len(levels):
        out["note"] = "no pull level in this traversal"
        return out
    top = max(pulls, key=lambda i: levels[i]["arcs"])
    unvisited = n - 1 - sum(l["discovered"] for l in levels[:top])
    b = unvisited * 12 + levels[top]["arcs"] * (4 + 1 / 8) + levels[top]["discovered"] * 4 + n / 8
    # time of the pull body alone (trace ids 32 = pull starts, 33 = stats posted) if the timeline has it, else the level
    t_ms, t_gather = level_ms[top], None
    if timeline:
        k = -1
        for j, (t, ident) in enumerate(timeline):
            if ident == 31:
                k += 1
                if k == pulls.index(top):
                    # (marks 35 / 36 -- CTA 0 out of chunks, grid barrier passed -- sit between 32 and 33)
                    t32 = next((tt for tt, ii in timeline[j + 1:j + 6] if ii == 32), None)
                    t33 = next((tt for tt, ii in timeline[j + 1:j + 6] if ii == 33), None)
                    if t32 is not None and t33 is not None and t33 > t32:
                        t_gather, t_ms = (t32 - t) * 1e-3, (t33 - t32) * 1e-3
                    break
    gbs = b / world / (t_ms * 1e-3) / 1e9
    out.update({"kernel": "p2p_pull_levels_kernel: bfs_pull_body of the heaviest pull level (per GPU: 1/P of the rows)",
                "achieved": gbs, "frac": gbs / peak, "algorithmic_bytes_per_launch": b / world, "launch_ms": t_ms,
                "bytes_model": "SURVEY.md 8d pull: |U|*12 + a*(4+1/8) + |F_next|*4 + n/8, divided by P", "level": top})
    if t_gather:
        nb = (world - 1) * (n // 8) // world
        out["nvlink"].update({"achieved_gbs": nb / (t_gather * 1e-3) / 1e9, "frac": nb / (t_gather * 1e-3) / 1e9 / 900.0,
                              "gather_ms": t_gather, "note": "the allgather of the heaviest pull level: peer loads of (P-1) bitmap slices"})
    return out


def run(args) -> int:
    import torch
    from bench import METRIC_MG
    import torch.distributed as dist
    import mini_b200 as mb
    from mini_b200 import dist as D
    from mini_b200.p2p import P2PBfs

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
    ctx.set_level_loop(mb.LOOP_HOST if getattr(args, "loop", "graph") == "host" else mb.LOOP_GRAPH)
    t0 = time.time()
    g = ctx.prepare_graph(D.build_rank_graph(ctx, scale, ef, 1, rank, world))
    build_s = time.time() - t0
    comm = D.TorchComm(ctx.torch_device)
    mode = "beamer" if args.mg_mode == "beamer" else "push"
    exchange = args.exchange

    if exchange == "p2p":
        rk = P2PBfs(ctx, rank, world, n, m, g)
        rk.connect_torch_distributed()
        rk.prepare(mode)
        dist.barrier()

        def run_bfs(md=mode):
            rk.run(0, md)
            return rk.levels
    else:
        rk = D.GpuRank(ctx, rank, world, n, g)
        drivers = {md: D.DistBFS(rk, comm, n, m, mode=md) for md in ("beamer", "push")}

        def run_bfs(md=mode):
            drivers[md].run(0)
            return drivers[md].levels

    sampler = None
    if rank == 0:
        from bench import ClockSampler
        sampler = ClockSampler(dev)
        sampler.start()

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

    for _ in range(args.warmup):
        run_bfs()
    ms_total, launches = timed(args.steps, run_bfs)
    if exchange == "p2p":
        rk.level_loop_timed = rk.level_loop
    levels = list(run_bfs())
    level_ms = [round(l.get("level_ms", 0.0), 4) for l in levels]
    timeline = None
    if exchange == "p2p":
        # one more run with the kernels' own timeline (outside the timed region; graph-driven loop: the levels run
        # inside persistent kernels, so the per-level times come from %globaltimer marks, not from CUDA events)
        rk.set_trace(True)
        run_bfs()
        rk.set_trace(False)
        timeline = rk.last_trace()
        lt = rk.level_times_ms()
        if len(lt) == len(levels):
            level_ms = [round(x, 4) for x in lt]
        else:                                   # host-driven loop: CUDA events per level
            rk.run(0, mode, timing=True)
            level_ms = [round(l["level_ms"], 4) for l in rk.levels]
    reached_arcs, reached_vertices = comm.all_reduce_sum([rk.reached_degree_sum(), rk.reached_count()])
    value = reached_arcs * args.steps / (ms_total * 1e-3) / 1e9
    props_ok, level_hist = D.verify_bfs_properties(rk, comm, 0)

    # end to end: every step uploads the rank's initial labels from pinned host memory, runs, downloads labels
    h_init = torch.full((rk.n_local,), -1, dtype=torch.int32).pin_memory()
    h_out = torch.empty(rk.n_local, dtype=torch.int32).pin_memory()

    def e2e_step():
        rk.labels.copy_(h_init, non_blocking=True)
        run_bfs()
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

    # the other traversal mode on the same partitioned graph (push-only when the headline is direction-optimising)
    other = "push" if mode == "beamer" else "beamer"
    other_line = None
    if args.mg_other:
        for _ in range(2):
            run_bfs(other)
        o_ms, _ = timed(max(2, args.steps // 2), lambda: run_bfs(other))
        o_ms /= max(2, args.steps // 2)
        o_levels = list(run_bfs(other))
        other_line = {"mode": other, "ms_per_step": o_ms, "value": reached_arcs / (o_ms * 1e-3) / 1e9, "unit": "GTEPS",
                      "sent_vertices": sum(l["sent"] for l in o_levels), "levels": len(o_levels)}

    all_launches = comm.all_reduce_sum([launches])[0]
    sent = sum(l["sent"] for l in levels)

    # the same graph on ONE GPU (rank 0 builds the whole scale-26 CSR: 8 GiB + offsets), same modes: the
    # denominator of "scaling from 1 to N GPUs on scale-26" (north_star); the other ranks wait at the barrier
    single = None
    lab1 = None
    oracle_parity = None
    if args.mg_single and rank == 0:
        try:
            g1 = ctx.prepare_graph(ctx.rmat_graph(scale, ef, 1))
            lab1 = torch.empty(g1.n, dtype=torch.int32, device=ctx.torch_device)
            single = {}
            for md, flag in (("beamer", mb.BFS_BEAMER), ("push", mb.BFS_PUSH)):
                for _ in range(3):
                    ctx.bfs(g1, 0, flag, 15.0, 18.0, labels=lab1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = max(3, args.steps // 2)
                e0.record()
                for _ in range(reps):
                    ctx.bfs(g1, 0, flag, 15.0, 18.0, labels=lab1)
                e1.record()
                torch.cuda.synchronize()
                ms1 = e0.elapsed_time(e1) / reps
                single[md] = {"ms_per_step": ms1, "value": reached_arcs / (ms1 * 1e-3) / 1e9, "unit": "GTEPS"}
            if getattr(args, "mg_oracle", True):
                # full-size parity, part 1: the single-GPU labels against the 64-bit CPU oracle BFS (bfs_problem.hxx:52-72
                # restated with int64 offsets -- the reference's int32 CSR cannot hold 2^31 arcs) on the device-built CSR
                import numpy as np
                import oracle
                t0 = time.time()
                o = oracle.CSR(g1.n, g1.offsets_host(), g1.col_indices.cpu().numpy())
                ref = oracle.bfs(o, 0)
                oracle_parity = {"single_gpu_labels_bit_exact_vs_cpu_oracle_int64": bool(np.array_equal(ref, lab1.cpu().numpy())),
                                 "oracle_s": round(time.time() - t0, 1)}
                del o, ref
            del g1
        except Exception as e:   # noqa: BLE001  (e.g. not enough free HBM next to the partition)
            single = {"unavailable": repr(e)[:200]}
            lab1 = None
    dist.barrier()
    # full-size parity, part 2: every rank's label slice against the single-GPU labels of the same graph, bit-exact
    slices_equal = None
    if args.mg_single:
        have = torch.tensor([1 if (rank != 0 or lab1 is not None) else 0], device=ctx.torch_device)
        dist.all_reduce(have, op=dist.ReduceOp.MIN)
        if int(have.item()) == 1:
            run_bfs(mode)
            if lab1 is None:
                lab1 = torch.empty(n, dtype=torch.int32, device=ctx.torch_device)
            dist.broadcast(lab1, src=0)
            from mini_b200 import partition as PT
            log_p = PT.log2_ranks(world)
            r = torch.arange(rk.n_local, dtype=torch.int64, device=ctx.torch_device)
            sw = (((r * PT.GOLDEN) & 0xFFFFFFFF) >> (32 - log_p)) if log_p else torch.zeros_like(r)
            gid = (r << log_p) | (rank ^ sw)
            eq = torch.tensor([int(torch.equal(lab1[gid], rk.labels))], device=ctx.torch_device)
            dist.all_reduce(eq, op=dist.ReduceOp.MIN)
            slices_equal = bool(int(eq.item()) == 1)
            del r, sw, gid
    lab1 = None
    torch.cuda.empty_cache()
    dist.barrier()

    if rank == 0:
        clocks = sampler.stop()
        ex_text = ("frontier exchange fused into the kernels over NVLink peer memory (advance flush stores into the owner's "
                   "inbox; gather+OR of bitmap slices; flag barriers) -- no NCCL on the data path") if exchange == "p2p" else \
                  "NCCL alltoallv (push) / allgather of bitmap slices (pull)"
        line = {
            "metric": METRIC_MG, "value": value, "unit": "GTEPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"direction-optimising BFS from vertex 0, RMAT scale-{scale} ef16 symmetrised "
                                   f"(n={n}, m={m}), cyclic 1D vertex partition over {world} GPUs, {ex_text}",
                       "mode": mode, "exchange": exchange,
                       "level_loop": getattr(rk, "level_loop_timed", "host"), "teps_numerator": "sum of deg(v) over reached v",
                       "reached_arcs": reached_arcs, "reached_vertices": reached_vertices,
                       "parallelism": f"1d-cyclic x{world}",
                       "l2": "inputs larger than L2 (per-rank col_indices %d MiB)" % (g.m * 4 >> 20),
                       "graph_build_s": round(build_s, 2), "levels": _level_rows(levels), "level_ms": level_ms},
            "clocks": clocks,
            "e2e": {"value": reached_arcs * args.steps / e2e_s / 1e9, "unit": "GTEPS",
                    "h2d_bytes_per_step": rk.n_local * 4 * world, "d2h_bytes_per_step": rk.n_local * 4 * world,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": all_launches,
            "roofline": _mg_roofline(levels, level_ms, timeline, n, world, sent),
            "cpu_baseline": None,
            "other_mode": other_line,
            "single_gpu_same_graph": single,
            "parity": {"bfs_properties_hold_at_full_scale": props_ok, "vertices_per_level": level_hist,
                       "every_rank_slice_bit_exact_vs_1gpu_labels": slices_equal},
        }
        if oracle_parity:
            line["parity"].update(oracle_parity)
        if single and mode in single:
            line["speedup_vs_1gpu_same_graph"] = single[mode]["ms_per_step"] / (ms_total / args.steps)
            line["scaling_efficiency_same_workload"] = line["speedup_vs_1gpu_same_graph"] / world
            if other_line and other in single:
                other_line["speedup_vs_1gpu_same_graph"] = single[other]["ms_per_step"] / other_line["ms_per_step"]
                other_line["scaling_efficiency_same_workload"] = other_line["speedup_vs_1gpu_same_graph"] / world
        print(json.dumps(line))
    dist.barrier()
    if exchange == "p2p":
        rk.close()
    ctx.close()
    dist.destroy_process_group()
    return 0 if (props_ok and slices_equal is not False) else 1
