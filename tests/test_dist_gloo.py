"""world_size-2 (and 4) gloo runs of the partitioned-BFS host logic on CPU: DistBFS + TorchComm with a
NumPy stand-in for the per-rank kernels (tests/numpy_rank.py), checked against the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,mode,src", [(2, "beamer", 0), (2, "push", 3), (4, "beamer", 0)])
def test_dist_bfs_over_gloo(world, mode, src):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests") + os.pathsep + ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + world * 7 + src), os.path.join(ROOT, "tests", "run_dist_bfs.py"),
           "--scale", "10", "--backend", "gloo", "--mode", mode, "--src", str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "DIST_BFS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    if mode == "beamer":
        assert "P" in r.stdout.split("dirs=")[1].split()[0]     # at least one pull level was taken
