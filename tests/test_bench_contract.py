"""CPU checks of bench.py's contract for the reference arm (the GPU arm needs a B200): ONE JSON line on stdout with
the agreed keys; under torchrun only rank 0 works and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, *args):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "14",
                           "--steps", "2", "--warmup", "1", *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bfs_gteps_rmat22_push" and d["unit"] == "GTEPS"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"] and "scale-14" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_names_the_multi_gpu_metric_for_n_gt_1():
    """N=1 (scale-22 push) and N>1 (scale-26 direction-optimising) are different workloads: the metric names differ so
    that nobody divides one by the other."""
    r = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["metric"] == "bfs_gteps_rmat26_do" and d["n_gpus"] == 2


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])


def test_multi_gpu_roofline_reads_the_pull_marks_of_the_kernel_timeline():
    """bench_multi._mg_roofline takes the pull body's time (marks 32 -> 33) and the allgather's (31 -> 32) from the
    kernels' own timeline; the marks 35 / 36 the pull-levels kernel logs in between must not hide them."""
    import bench_multi as bm
    levels = [dict(direction="push", frontier=1, arcs=10, discovered=5, sent=0),
              dict(direction="pull", frontier=5, arcs=1000, discovered=50, sent=0)]
    tl = [(0.0, 20), (10.0, 64), (11.0, 30), (12.0, 31), (30.0, 32), (100.0, 35), (140.0, 36), (142.0, 33), (143.0, 34), (144.0, 65)]
    r = bm._mg_roofline(levels, [0.01, 0.134], tl, 1 << 20, 8, 0)
    assert abs(r["launch_ms"] - 0.112) < 1e-9 and abs(r["nvlink"]["gather_ms"] - 0.018) < 1e-9
    assert r["frac"] is not None and r["nvlink"]["frac"] is not None
    r = bm._mg_roofline(levels, [0.01, 0.134], [(0.0, 20)], 1 << 20, 8, 0)      # no marks: the whole level's time
    assert abs(r["launch_ms"] - 0.134) < 1e-9 and r["nvlink"]["frac"] is None
