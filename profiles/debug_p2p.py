"""Debug aid: one virtual-rank configuration of tests/test_gpu_multi.py with the per-level records of every rank and the
first mismatching vertices against the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from mini_b200 import partition as PT
import test_gpu_multi as T

scale, ef, seed, world, src, mode, small = 14, 8, 3, 4, 77, "beamer", "small-tiny"
if len(sys.argv) > 1:
    scale, ef, seed, world, src = [int(x) for x in sys.argv[1:6]]
    mode, small = sys.argv[6], sys.argv[7]
o = oracle.rmat_csr(scale, ef, seed)
ref = oracle.bfs(o, src)
print("oracle level histogram:", np.bincount(ref[ref >= 0]).tolist())
for attempt in range(3):
    res = T._virtual_ranks_p2p(scale, ef, seed, world, src, mode, small=small)
    labels = np.empty(1 << scale, np.int32)
    for r in range(world):
        labels[PT.global_ids(r, world, (1 << scale) // world)] = res[r][0]
    bad = np.nonzero(labels != ref)[0]
    print(f"attempt {attempt}: mismatches {len(bad)}")
    for l in res[0][2]:
        print("   ", {k: l[k] for k in ("direction", "exchange", "frontier", "arcs", "discovered", "sent")})
    for v in bad[:12]:
        print(f"    v={v} owner={int(PT.owner(v, world))} got={labels[v]} want={ref[v]} deg={o.offsets[v+1]-o.offsets[v]}")
