"""Generates tests/golden/*.json by running the REFERENCE's own CPU code
(oracle/_ref/libref_cpu.so = /root/reference headers compiled unmodified, see
oracle/ref_shim.cu) on the reference's own fixtures and on small seeded RMAT
graphs.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The outputs are committed; tests read them and never touch /root/reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

REF = "/root/reference/gunrock/tests"


def mtx_edges(path):
    with open(path) as f:
        lines = [ln.split() for ln in f.read().splitlines() if ln.strip() and not ln.startswith("%")]
    return [int(x) for x in lines[0][:3]], [[float(x) if k == 2 else int(x) for k, x in enumerate(t)] for t in lines[1:]]


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", name)


def fixture(algo, undirected, src=0, mtx="test.mtx"):
    path = f"{REF}/{algo}/{mtx}"
    header, edges = mtx_edges(path)
    g, csc_eq = oracle.ref_load_graph(path, undirected)
    rec = {
        "source": f"gunrock/tests/{algo}/{mtx} via load_graph(file,{str(undirected).lower()},false) graph.hxx:96-223",
        "mtx_header": header, "mtx_edges": edges, "undirected": undirected, "src": src,
        "n": g.n, "m": g.m, "offsets": g.offsets.tolist(), "indices": g.indices.tolist(),
        "weights": g.weights.tolist(), "csc_equals_csr": csc_eq,
    }
    labels, _ = oracle.ref_bfs(g, src)
    rec["bfs_labels"] = labels.tolist()          # bfs_problem.hxx:52-72
    preds, _ = oracle.ref_sssp_preds(g, src)
    rec["sssp_preds"] = preds.tolist()           # sssp_problem.hxx:59-88
    return rec


def rmat_case(scale, ef, seed, src=0):
    g = oracle.rmat_csr(scale, ef, seed, weighted=True)
    labels, _ = oracle.ref_bfs(g, src)
    preds, _ = oracle.ref_sssp_preds(g, src)
    import hashlib
    return {
        "source": "RMAT (oracle.c generator) -> reference bfs_problem_t::cpu / sssp_problem_t::cpu",
        "scale": scale, "edge_factor": ef, "seed": seed, "wseed": 7, "src": src, "n": g.n, "m": g.m,
        "csr_sha256": hashlib.sha256(g.offsets.tobytes() + g.indices.tobytes() + g.weights.tobytes()).hexdigest(),
        "offsets_head": g.offsets[:9].tolist(), "indices_head": g.indices[:16].tolist(),
        "weights_head": g.weights[:16].tolist(),
        "bfs_labels_sha256": hashlib.sha256(labels.tobytes()).hexdigest(),
        "bfs_labels_hist": np.bincount(labels + 1).tolist(),    # index 0 = unreached
        "bfs_labels": labels.tolist() if g.n <= 1024 else None,
        "sssp_preds_sha256": hashlib.sha256(preds.tobytes()).hexdigest(),
        "sssp_preds": preds.tolist() if g.n <= 1024 else None,
    }


if __name__ == "__main__":
    assert oracle.have_ref(), "needs /root/reference to build oracle/_ref"
    dump("ref_fixture_bfs.json", fixture("bfs", True))
    dump("ref_fixture_sssp_directed.json", fixture("sssp", False))
    dump("ref_fixture_sssp_undirected.json", fixture("sssp", True))
    dump("ref_fixture_pr.json", fixture("pr", True))
    dump("ref_fixture_kcore.json", fixture("kcore", True, mtx="test_kcore.mtx"))
    dump("ref_fixture_coloring.json", fixture("coloring", True))
    dump("ref_rmat_s8.json", rmat_case(8, 16, 1))
    dump("ref_rmat_s10.json", rmat_case(10, 16, 1))
    dump("ref_rmat_s16.json", rmat_case(16, 16, 1))
    dump("ref_rmat_s12_seed3.json", rmat_case(12, 8, 3, src=5))
