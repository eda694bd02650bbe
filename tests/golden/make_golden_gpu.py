"""Generates tests/golden/ref_gpu_*.json by running the REFERENCE's own GPU primitives
(oracle/_ref/ref_gpu_{bfs,sssp,pr} = /root/reference headers compiled unmodified for sm_100, see
oracle/ref_gpu_driver.cu) on the reference's fixtures and on small seeded RMAT graphs.

Needs a GPU, so it runs on the B200 box (the binaries are built in the container, where
/root/reference lives, and travel with the snapshot):

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/golden'

The outputs are then copied into tests/golden/ and committed; tests read them and never touch
/root/reference.  These vectors pin what the reference's own tests do not validate: the
neighbourhood reduce (neighborhood.hxx:12-70) and the PR driver (pr_enactor.hxx:41-79), and the
SSSP *distances* (test_sssp.cu compares the racy preds only).
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def custom_values(n):
    """Non-uniform initial ranks (the PR driver itself keeps all ranks equal on a symmetric graph)."""
    v = np.arange(n, dtype=np.uint64)
    return (0.15 + ((v * np.uint64(2654435761)) % np.uint64(1000)).astype(np.float64) / 1000.0).astype(np.float32)


def lens_of(stdout):
    return [int(x) for x in re.findall(r"finished iteration:\d+ output length: (\d+)", stdout)]


def pr_cases(g, tag):
    rec = {"source": "pr_enactor_t::enact (pr_enactor.hxx:41-79) = neighborhood_kernel<plus_t<float>> + filter_kernel, "
                     "reference GPU code compiled unmodified for sm_100, run on a B200", "graph": tag, "n": g.n, "m": g.m,
           "values_formula": "0.15 + ((v * 2654435761) % 1000) / 1000 as float32", "cases": []}
    for max_iter, custom in ((1, False), (10, False), (1, True), (3, True), (10, True)):
        vals = custom_values(g.n) if custom else None
        out, secs = oracle.ref_gpu("pr", g, max_iter=max_iter, values=vals)
        c = {"max_iter": max_iter, "custom_values": custom, "frontier_lens": lens_of(out["stdout"])}
        if g.n <= 64:
            c["current"] = [float(x) for x in out["current"]]
            c["reduced"] = [float(x) for x in out["reduced"]]
        else:
            import base64
            c["current_f32_b64"] = base64.b64encode(out["current"].astype(np.float32).tobytes()).decode()
            c["reduced_f32_b64"] = base64.b64encode(out["reduced"].astype(np.float32).tobytes()).decode()
        rec["cases"].append(c)
    if g.n > 64:
        rec["encoding"] = "current / reduced: base64 of little-endian float32[n]"
    return rec


def traversal_cases(g, tag, srcs):
    rec = {"source": "bfs_enactor_t::enact_pushpull (bfs_enactor.hxx:41-117) / sssp_enactor_t::enact (sssp_enactor.hxx:40-72), "
                     "reference GPU code compiled unmodified for sm_100, run on a B200", "graph": tag, "n": g.n, "m": g.m,
           "cases": []}
    import hashlib
    for src in srcs:
        b, _ = oracle.ref_gpu("bfs", g, src=src)
        b2, _ = oracle.ref_gpu("bfs", g, src=src, alpha=2.0)     # forces the push -> pull switch
        s, _ = oracle.ref_gpu("sssp", g, src=src, queue_sizing=4.0)
        c = {"src": src, "bfs_labels_sha256": hashlib.sha256(b["labels"].tobytes()).hexdigest(),
             "bfs_pushpull_labels_sha256": hashlib.sha256(b2["labels"].tobytes()).hexdigest(),
             "sssp_dist_sha256": hashlib.sha256(s["labels"].tobytes()).hexdigest()}
        if g.n <= 64:
            c["bfs_labels"] = b["labels"].tolist()
            c["bfs_pushpull_labels"] = b2["labels"].tolist()
            c["sssp_dist"] = [float(x) for x in s["labels"]]
        rec["cases"].append(c)
    return rec


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    assert oracle.have_ref_gpu(), "oracle/_ref/ref_gpu_* missing: run `make -C oracle` where /root/reference exists"

    def dump(name, obj):
        with open(os.path.join(outdir, name), "w") as f:
            json.dump(obj, f, separators=(",", ":"))
        print("wrote", name)

    def fixture(name):
        rec = json.load(open(os.path.join(HERE, name)))
        return oracle.CSR(rec["n"], rec["offsets"], rec["indices"], rec["weights"])

    gpr = fixture("ref_fixture_pr.json")
    dump("ref_gpu_pr_fixture.json", pr_cases(gpr, "ref_fixture_pr.json"))
    g12 = oracle.rmat_csr(12, 16, 1, weighted=True)
    dump("ref_gpu_pr_rmat_s12.json", pr_cases(g12, "rmat scale 12 ef 16 seed 1"))
    dump("ref_gpu_traversal_fixture_bfs.json", traversal_cases(fixture("ref_fixture_bfs.json"), "ref_fixture_bfs.json", [0, 3]))
    dump("ref_gpu_traversal_fixture_sssp.json",
         traversal_cases(fixture("ref_fixture_sssp_undirected.json"), "ref_fixture_sssp_undirected.json", [0, 2]))
    dump("ref_gpu_traversal_rmat_s12.json", traversal_cases(g12, "rmat scale 12 ef 16 seed 1 wseed 7", [0, 5]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_gpu_out"))
