#!/usr/bin/env python
"""bench.py -- BFS GTEPS on synthetic RMAT graphs (BASELINE.json metric), with the advance
kernel's HBM roofline and the reference's CPU BFS timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale S]

A "step" is one BFS from vertex 0 over the whole graph (source in frontier -> labels final).
N = 1: BASELINE.json configs[1], RMAT scale-22 ef16, push BFS (LB advance + fused uniquify filter).
N > 1: configs[3], RMAT scale-26 ef16, cyclic 1D vertex partition, one rank per GPU, frontier exchange
fused into the kernels over NVLink peer memory (mini_b200.p2p; --exchange nccl = the NCCL form,
mini_b200.dist).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bfs_gteps_rmat"
UNIT = "GTEPS"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


WORK_CREATE = True    # --advance quad (default): the advance launch also produces the next level's scan


def push_level_bytes(level, offset_bytes=4):
    """Algorithmic bytes of one push level: |F|(4+2*O) + m_F*(4+4) + |F_next|*(4+4); the work-creating
    advance additionally reads two offsets and writes row bounds + scan position per emitted vertex (the scan
    kernel's bytes, moved into this launch)."""
    b = level["frontier_len"] * (4 + 2 * offset_bytes) + level["arcs"] * 8 + level["discovered"] * 8
    if WORK_CREATE:
        b += level["discovered"] * (2 * offset_bytes + 8 + 4)
    return b


def sssp_level_bytes(level, offset_bytes=4):
    """SURVEY.md 8d, SSSP advance: |F|(4+2*O+4) + m_l*(4 idx + 4 weight + 4 label) + improved*(4+4)."""
    return level["frontier_len"] * (8 + 2 * offset_bytes) + level["arcs"] * 12 + level["discovered"] * 8


def reduce_level_bytes(level, offset_bytes=4):
    """SURVEY.md 8d, neighborhood_reduce fp32: |F|(4+2*O) + m_F*(4 idx + 4 gather) + |F|*4."""
    return level["frontier_len"] * (4 + 2 * offset_bytes) + level["arcs"] * 8 + level["frontier_len"] * 4


def _per_level(ctx_call, reps, bytes_fn, peak):
    """Runs ctx_call(timing=True) reps times; per-level advance_ms averaged; roofline of the heaviest launch."""
    runs = [ctx_call().levels for _ in range(reps)]
    # (a racy fixed-point iteration such as SSSP may take a different number of iterations from run to run:
    # average the runs that have the same shape as the first one)
    lv = runs[0]
    same = [r for r in runs if len(r) == len(lv)]
    lv_ms = [sum(r[i]["advance_ms"] for r in same) / len(same) for i in range(len(lv))]
    top = max(range(len(lv)), key=lambda i: lv[i]["arcs"])
    b = bytes_fn(lv[top])
    gbs = b / (lv_ms[top] * 1e-3) / 1e9
    tot_b, tot_ms = sum(bytes_fn(l) for l in lv), sum(lv_ms)
    return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
            "algorithmic_bytes_per_launch": b, "launch_ms": lv_ms[top], "arcs_per_launch": lv[top]["arcs"],
            "all_launches": {"achieved": tot_b / (tot_ms * 1e-3) / 1e9, "frac": tot_b / (tot_ms * 1e-3) / 1e9 / peak,
                             "kernel_ms_sum": tot_ms, "launches": len(lv)}}


def run_sssp_leg(ctx, mb, args, peak):
    """BASELINE.json configs[2]: SSSP, idempotent LB advance, RMAT scale-22 ef16, integer weights 1..64 (fp32)."""
    import torch
    scale = args.scale or 22
    g = ctx.rmat_graph(scale, 16, 1, weighted=True)
    dist = torch.empty(g.n, dtype=torch.float32, device=ctx.torch_device)
    steps = max(3, min(args.steps, 20))
    for _ in range(3):
        ctx.sssp(g, 0, dist=dist)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(steps):
        _, st = ctx.sssp(g, 0, dist=dist)
        launches += st.launches
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    off = g.row_offsets.to(torch.int64) & 0xFFFFFFFF
    reached_arcs = int((off[1:] - off[:-1])[dist < 3.0e38].sum().item())
    roof = _per_level(lambda: ctx.sssp(g, 0, dist=dist, timing=True)[1], 5, sssp_level_bytes, peak)
    roof["kernel"] = "quad_advance_kernel<SsspRelaxQ,COMPACT> (heaviest iteration)"
    return {"workload": f"SSSP from vertex 0, RMAT scale-{scale} ef16 symmetrised, uniform integer weights [1,64] "
                        "as fp32, idempotent LB advance (Bellman-Ford frontier iterations)",
            "value": reached_arcs / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "iterations": st.num_levels, "relaxed_arcs": st.total_arcs, "reached_arcs": reached_arcs,
            "gpu_launches": launches, "roofline": roof}


def run_reduce_leg(ctx, mb, args, peak):
    """BASELINE.json configs[4]: neighborhood_reduce PageRank-style fp32 pull-sum over dynamic frontiers,
    RMAT scale-24 ef16, max_iter 10 (pr_enactor.hxx:41-79)."""
    import torch
    scale = args.reduce_scale
    g = ctx.rmat_graph(scale, 16, 1)
    for _ in range(2):
        ctx.pr(g, 10, False)
    torch.cuda.synchronize()
    steps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = 0
    for _ in range(steps):
        _, _, lens, st = ctx.pr(g, 10, False)
        launches += st.launches
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    roof = _per_level(lambda: ctx.pr(g, 10, False, timing=True)[3], 3, reduce_level_bytes, peak)
    roof["kernel"] = "quad_segreduce_kernel<float,PlusF32> (heaviest iteration: frontier = all vertices)"
    roof["traffic"] = None
    tp = os.path.join(ROOT, "profiles", "advance_traffic.json")
    if scale == 24 and os.path.exists(tp):
        try:
            roof["traffic"] = json.load(open(tp)).get("scale24_reduce_top_launch_dram_bytes")
        except Exception:
            pass
    return {"workload": f"PageRank-style neighborhood_reduce fp32 pull-sum, RMAT scale-{scale} ef16 symmetrised "
                        f"(n={g.n}, m={g.m}), 10 iterations over the filter's dynamic frontiers",
            "value": st.total_arcs / (ms * 1e-3) / 1e9, "unit": "G arcs reduced/s", "ms_per_step": ms, "steps": steps,
            "iterations": st.num_levels, "frontier_lens": lens, "reduced_arcs": st.total_arcs,
            "gpu_launches": launches, "roofline": roof}


def cpu_bfs_baseline(off64, idx, runs):
    """The reference's CPU validation BFS (bfs_problem.hxx:52-72) on this box's host cores:
    the unmodified reference code if oracle/_ref was built, else the oracle port."""
    import numpy as np
    import oracle
    g = oracle.CSR(len(off64) - 1, off64, idx)
    kind = "reference" if (oracle.have_ref() and g.m < 2 ** 31) else "port"
    times, labels = [], None
    for _ in range(runs):
        if kind == "reference":
            labels, t = oracle.ref_bfs(g, 0)
        else:
            labels, t = oracle.bfs_timed(g, 0)
        times.append(t)
    reached = oracle.reached_arcs(g, labels)
    return kind, times, reached, labels


def run_reference_arm(args):
    """--impl reference: the reference's own CPU BFS (serial; it has no threaded CPU code) on the same
    workload, each step one full BFS from vertex 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    # N = 1: the arm's own workload (scale 22).  N > 1: the arm runs scale 26, whose CSR (2^31 arcs) the reference's
    # int32 CPU path cannot hold and whose host-side build alone takes minutes: the bounded sample is the same
    # generator at scale 24 (a quarter of the vertices and arcs; GTEPS is a rate)
    scale = args.scale or (22 if args.gpus <= 1 else 24)
    t0 = time.time()
    g = oracle.rmat_csr(scale, 16, 1)
    gen_s = time.time() - t0
    for _ in range(args.warmup):
        cpu_bfs_baseline(g.offsets, g.indices, 1)
    kind, times, reached, _ = cpu_bfs_baseline(g.offsets, g.indices, args.steps)
    total = sum(times)
    val = reached * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"BFS from vertex 0, RMAT scale-{scale} ef16 symmetrised (n={g.n}, m={g.m}), "
                               "reference CPU validation BFS bfs_problem.hxx:52-72"
                               + ("" if args.gpus <= 1 or args.scale else " -- bounded sample of the N>1 arm's scale-26 workload"),
                   "graph_build_s": round(gen_s, 2)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{len(times)} full BFS runs from vertex 0 on the whole scale-{scale} graph",
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_single_gpu(args):
    import numpy as np
    import torch
    import mini_b200 as mb

    scale = args.scale or 22
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    ctx = mb.Context(dev)
    global WORK_CREATE
    WORK_CREATE = args.advance == "quad" and (args.scale or 22) <= 23   # engine.cuh WORK_CREATE_MAX_N
    ctx.set_advance_impl({"lbs": mb.ADVANCE_LBS, "rescan": mb.ADVANCE_QUAD_RESCAN}.get(args.advance, mb.ADVANCE_QUAD))
    ctx.set_level_loop(mb.LOOP_HOST if args.loop == "host" else mb.LOOP_GRAPH)
    g = ctx.prepare_graph(ctx.rmat_graph(scale, 16, 1))   # graph build + one-time derived data: outside the timed region
    mode = {"push": mb.BFS_PUSH, "beamer": mb.BFS_BEAMER}[args.mode]
    sampler = ClockSampler(dev)
    sampler.start()

    # ---- device-resident throughput: W warm-up + K timed steps, CUDA events on the launching stream
    labels = torch.empty(g.n, dtype=torch.int32, device=ctx.torch_device)
    for _ in range(args.warmup):
        ctx.bfs(g, 0, mode, 15.0, 18.0, labels=labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        _, st = ctx.bfs(g, 0, mode, 15.0, 18.0, labels=labels)
        launches += st.launches
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    loop_used = st.level_loop
    reached_arcs = g.degrees_sum_reached(labels)
    value = reached_arcs * args.steps / (ms_total * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (lbs_advance_kernel<BfsPushOp>): per-launch CUDA-event times
    peak, peak_src = _peaks()
    reps = max(3, min(args.steps, 20))
    lv_ms, lv = None, None
    for _ in range(reps):
        _, st = ctx.bfs(g, 0, mb.BFS_PUSH, labels=labels, timing=True)
        if lv_ms is None:
            lv, lv_ms = st.levels, [0.0] * len(st.levels)
        for i, l in enumerate(st.levels):
            lv_ms[i] += l["advance_ms"] / reps
    tot_bytes = sum(push_level_bytes(l) for l in lv)
    tot_ms = sum(lv_ms)
    top = max(range(len(lv)), key=lambda i: lv[i]["arcs"])
    top_gbs = push_level_bytes(lv[top]) / (lv_ms[top] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "advance_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"scale{scale}_top_level_dram_bytes")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": ("quad_advance_kernel<BfsPushQ,COMPACT>" if args.advance != "lbs" else "lbs_advance_kernel<BfsPushOp,COMPACT>") + " (heaviest BFS level)",
        "achieved": top_gbs, "peak": peak, "unit": "GB/s", "frac": top_gbs / peak, "traffic": traffic,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": push_level_bytes(lv[top]), "launch_ms": lv_ms[top],
        "arcs_per_launch": lv[top]["arcs"],
        "all_levels": {"achieved": tot_bytes / (tot_ms * 1e-3) / 1e9, "frac": tot_bytes / (tot_ms * 1e-3) / 1e9 / peak,
                       "advance_ms_sum": tot_ms,
                       "levels": [dict(frontier=l["frontier_len"], arcs=l["arcs"], discovered=l["discovered"],
                                       advance_ms=round(ms, 4)) for l, ms in zip(lv, lv_ms)]},
    }

    # ---- end to end through the host-buffer C-ABI call: H2D initial labels + BFS + D2H labels every step
    off_h = g.row_offsets.cpu().numpy().view(np.uint32)
    idx_h = g.col_indices.cpu().numpy()
    hg = ctx.host_graph_upload(off_h, idx_h)
    h_init = torch.full((g.n,), -1, dtype=torch.int32).pin_memory()
    h_init[0] = 0
    h_out = torch.empty(g.n, dtype=torch.int32).pin_memory()
    for _ in range(args.warmup):
        ctx.bfs_host(hg, 0, h_init.data_ptr(), h_out.data_ptr(), mode, 15.0, 18.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.bfs_host(hg, 0, h_init.data_ptr(), h_out.data_ptr(), mode, 15.0, 18.0)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert torch.equal(h_out, labels.cpu()), "e2e labels differ from the device-resident run"
    e2e = {"value": reached_arcs * args.steps / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": g.n * 4,
           "d2h_bytes_per_step": g.n * 4, "ms_per_step": 1e3 * e2e_s / args.steps,
           "note": "b200_bfs_host: bfs_problem_t ctor upload of the initial labels + enact + extract(); "
                   "graph resident as in test_bfs.cu (graph_to_device precedes the timer)"}
    ctx.host_graph_free(hg)
    clocks = sampler.stop()

    # ---- the reference's CPU BFS on this box, bounded sample, checked against the GPU labels
    if args.cpu_runs > 0:
        off64 = off_h.astype(np.int64)
        kind, times, cpu_reached, cpu_labels = cpu_bfs_baseline(off64, idx_h, args.cpu_runs)
        parity = bool(np.array_equal(cpu_labels, h_out.numpy()))
        cpu = {"value": cpu_reached * len(times) / sum(times) / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"{len(times)} full BFS runs from vertex 0 on the whole scale-{scale} graph "
                         f"({sum(times):.1f} s of CPU work)", "host_cores_available": os.cpu_count(),
               "labels_match_gpu": parity}
    else:
        parity, cpu = None, None

    # ---- the other single-GPU configurations of BASELINE.json (own graphs; the BFS graph is released first)
    gn, gm = g.n, g.m
    del g, labels
    torch.cuda.empty_cache()
    extras = {}
    if "sssp" in args.extras:
        extras["sssp"] = run_sssp_leg(ctx, mb, args, peak)
        torch.cuda.empty_cache()
    if "reduce" in args.extras:
        extras["neighborhood_reduce"] = run_reduce_leg(ctx, mb, args, peak)
        torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"BFS from vertex 0, RMAT scale-{scale} ef16 symmetrised (n={gn}, m={gm}), "
                               f"{args.mode} (LB advance + fused uniquify filter)",
                   "advance": args.advance + (" (work-creating: the flush writes the next level's row bounds and scan "
                                               "positions; their 20 B per emitted vertex are counted in the roofline's "
                                               "algorithmic bytes)" if args.advance == "quad" else ""),
                   "teps_numerator": "sum of deg(v) over reached v", "reached_arcs": reached_arcs,
                   "l2": "inputs larger than L2 (col_indices alone is %d MiB vs 126 MB L2)" % (gm * 4 >> 20),
                   "level_loop": loop_used + (" (one CUDA graph per BFS: WHILE/IF/SWITCH conditional nodes set by a "
                                                  "device-side decide kernel; one host sync per BFS)"
                                                  if loop_used == "graph" else " (one counter read-back per level)"),
                   "parallelism": "1 GPU"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "parity": {"bfs_labels_bit_exact_vs_cpu": parity},
    }
    line.update(extras)
    print(json.dumps(line))
    ctx.close()
    return 0 if parity in (True, None) else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=0)
    ap.add_argument("--mode", default="push", choices=["push", "beamer"])
    ap.add_argument("--cpu-runs", type=int, default=3)
    ap.add_argument("--advance", default="quad", choices=["quad", "rescan", "lbs"],
                    help="push-advance kernel: quad_advance.cuh creating the next level's scan in its flush (default), "
                         "the same with a scan kernel before every level, or the first-generation advance.cuh")
    ap.add_argument("--loop", default="graph", choices=["graph", "host"],
                    help="N=1 BFS level loop: one CUDA graph with device-side decisions, or host-driven")
    ap.add_argument("--extras", default="sssp,reduce",
                    help="N=1: extra legs after the BFS line: sssp (configs[2]) and/or reduce (configs[4]); '' = none")
    ap.add_argument("--reduce-scale", dest="reduce_scale", type=int, default=24)
    ap.add_argument("--mg-mode", dest="mg_mode", default="beamer", choices=["push", "beamer"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1 frontier exchange: fused into the kernels over NVLink peer memory, or NCCL collectives")
    ap.add_argument("--no-mg-other", dest="mg_other", action="store_false", help="N>1: skip the other traversal mode")
    ap.add_argument("--no-mg-single", dest="mg_single", action="store_false",
                    help="N>1: skip the same-graph single-GPU run on rank 0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from mini_b200 import dist_bench
        return dist_bench.run(args)
    return run_single_gpu(args)


def _main_with_clean_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun notices) also
    write to fd 1, so everything but our final line is diverted to stderr."""
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import io
    buf = io.StringIO()
    old, sys.stdout = sys.stdout, buf
    try:
        rc = main()
    finally:
        sys.stdout = old
        lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
        other = [l for l in buf.getvalue().splitlines() if not l.startswith("{")]
        for l in other:
            print(l, file=sys.stderr)
        if lines:
            os.write(real_stdout, (lines[-1] + "\n").encode())
        os.close(real_stdout)
    return rc


if __name__ == "__main__":
    sys.exit(_main_with_clean_stdout())
