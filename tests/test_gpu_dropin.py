"""Drop-in acceptance on the GPU box.
 * build/dropin/ref_test_{bfs,sssp,pr}: the reference's UNMODIFIED tests/<algo>/test_<algo>.cu compiled
   (in the build container, examples/Makefile) against include/gunrock instead of gunrock/src + moderngpu.
   They self-validate against the reference's own cpu() code and print "Correct." (test_bfs.cu:44-52).
 * build/dropin/frontier_driver: our driver on the same headers, operator path and engine path, RMAT input.
"""
import os
import subprocess

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "dropin")


def _run(exe, *args, timeout=300):
    r = subprocess.run([os.path.join(BIN, exe), *args], capture_output=True, text=True, timeout=timeout)
    return r.returncode, r.stdout + r.stderr


def _mtx(tmp_path, rec, name="g.mtx"):
    p = tmp_path / name
    with open(p, "w") as f:
        f.write(" ".join(str(x) for x in rec["mtx_header"]) + "\n")
        for e in rec["mtx_edges"]:
            f.write(" ".join(str(int(x)) if k < 2 else repr(x) for k, x in enumerate(e)) + "\n")
    return str(p)


@pytest.fixture(scope="module", autouse=True)
def binaries():
    if not os.path.exists(os.path.join(BIN, "frontier_driver")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)
    assert os.path.exists(os.path.join(BIN, "frontier_driver"))


def _need(exe):
    if not os.path.exists(os.path.join(BIN, exe)):
        pytest.skip(f"{exe} was not prebuilt (needs the reference tree at build time)")


@pytest.mark.parametrize("extra", [[], ["--alpha=2"], ["--alpha=100"], ["--src=3"], ["--src=6", "--alpha=1"]])
def test_unmodified_reference_test_bfs(tmp_path, golden, extra):
    _need("ref_test_bfs")
    rc, out = _run("ref_test_bfs", f"--file={_mtx(tmp_path, golden('ref_fixture_bfs.json'))}", *extra)
    assert rc == 0 and "Correct." in out and "Validation Error" not in out, out
    assert "pushed iterations:" in out and "elapsed time:" in out


def test_unmodified_reference_test_sssp(tmp_path, golden):
    _need("ref_test_sssp")
    for rec_name, flags in (("ref_fixture_sssp_directed.json", []), ("ref_fixture_sssp_undirected.json", ["--undirected"])):
        rc, out = _run("ref_test_sssp", f"--file={_mtx(tmp_path, golden(rec_name))}", "--queue-sizing=1.5", *flags)
        assert rc == 0 and "elapsed time:" in out, out
        # the reference's test validates the racy predecessor array: it prints one of the two verdicts,
        # and which one depends on who wrote preds last -- so the DISTANCES are asserted below instead, through our driver
        # on the same file, bit-exact against the oracle
        assert ("Correct," in out) or ("Validation Error." in out), out
        import numpy as np
        import oracle
        rec = golden(rec_name)
        out_bin = tmp_path / "dist.bin"
        rc, log = _run("frontier_driver", "--algo=sssp", f"--file={_mtx(tmp_path, rec)}", "--queue-sizing=1.5", *flags,
                       f"--dump={out_bin}")
        assert rc == 0 and "Correct." in log, log
        o = oracle.CSR(rec["n"], rec["offsets"], rec["indices"], rec["weights"])
        assert np.fromfile(out_bin, np.float32).tobytes() == oracle.sssp_dist(o, 0).tobytes()


def test_unmodified_reference_test_pr(tmp_path, golden):
    _need("ref_test_pr")
    rc, out = _run("ref_test_pr", f"--file={_mtx(tmp_path, golden('ref_fixture_pr.json'))}", "--max_iter=5")
    assert rc == 0 and "finished iteration:0 output length:" in out and "elapsed time:" in out, out


# ---- kcore and coloring as clients of the operator API (advance<has_output=false>, integer min / max
# neighbourhood reductions).  The reference's own headers do not compile for these two; its UNMODIFIED tests do,
# against include/gunrock.
def _dumped(tmp_path, *args, dtype="int32"):
    import numpy as np
    out = tmp_path / "result.bin"
    rc, log = _run("frontier_driver", *args, f"--dump={out}")
    assert rc == 0 and "Correct." in log, log
    return np.fromfile(out, dtype=dtype), log


def test_unmodified_reference_test_kcore(tmp_path, golden):
    _need("ref_test_kcore")
    rc, out = _run("ref_test_kcore", f"--file={_mtx(tmp_path, golden('ref_fixture_kcore.json'))}")
    assert rc == 0 and "Correct." in out and "Validation Error" not in out and "largest k-core: 6" in out, out


def test_unmodified_reference_test_coloring(tmp_path, golden):
    _need("ref_test_coloring")
    rec = golden("ref_fixture_coloring.json")
    rc, out = _run("ref_test_coloring", f"--file={_mtx(tmp_path, rec)}")
    assert rc == 0 and "elapsed time:" in out, out
    import numpy as np
    import oracle
    colors = [int(x) for x in out.strip().splitlines()[-1].split()]     # test_coloring.cu:41 prints d_colors
    want, _ = oracle.coloring(oracle.CSR(rec["n"], rec["offsets"], rec["indices"], rec["weights"]), 15485863, 10)
    assert colors == want.tolist()


@pytest.mark.parametrize("graph", ["fixture", "rmat8", "rmat11"])
def test_kcore_client_vs_oracle(tmp_path, golden, graph):
    """Core numbers bit-exact against the restated reference peel (oracle.kcore; kcore_problem.hxx:54-105)."""
    import numpy as np
    import oracle
    if graph == "fixture":
        rec = golden("ref_fixture_kcore.json")
        o = oracle.CSR(rec["n"], rec["offsets"], rec["indices"], rec["weights"])
        cores, log = _dumped(tmp_path, "--algo=kcore", f"--file={_mtx(tmp_path, rec)}")
    else:
        scale = int(graph[4:])
        o = oracle.rmat_csr(scale, 16, 1)
        cores, log = _dumped(tmp_path, "--algo=kcore", f"--rmat-scale={scale}")
    want, largest = oracle.kcore(o)
    assert np.array_equal(cores, want) and f"largest k-core: {largest}" in log


@pytest.mark.parametrize("graph,iters", [("fixture", 10), ("rmat10", 10), ("rmat14", 3)])
def test_coloring_client_vs_oracle(tmp_path, golden, graph, iters):
    """Colours bit-exact against the numpy restatement fed the same std::mt19937 hash stream (oracle.coloring);
    the driver itself checks that no arc joins two vertices of one colour."""
    import numpy as np
    import oracle
    if graph == "fixture":
        rec = golden("ref_fixture_coloring.json")
        o = oracle.CSR(rec["n"], rec["offsets"], rec["indices"], rec["weights"])
        colors, log = _dumped(tmp_path, "--algo=coloring", f"--file={_mtx(tmp_path, rec)}", f"--max_iter={iters}")
    else:
        scale = int(graph[4:])
        o = oracle.rmat_csr(scale, 16, 1)
        colors, log = _dumped(tmp_path, "--algo=coloring", f"--rmat-scale={scale}", f"--max_iter={iters}")
    want, lens = oracle.coloring(o, 15485863, iters)
    assert np.array_equal(colors, want)
    assert "uncoloured after each iteration: " + " ".join(str(x) for x in lens) in log


@pytest.mark.parametrize("args", [
    ["--algo=bfs", "--rmat-scale=14"],
    ["--algo=bfs", "--rmat-scale=14", "--alpha=2"],                 # operator-level push -> pull hand-over
    ["--algo=bfs", "--rmat-scale=15", "--src=5", "--alpha=0.5"],
    ["--algo=bfs", "--rmat-scale=16", "--builtin", "--mode=0"],
    ["--algo=bfs", "--rmat-scale=16", "--builtin", "--mode=2", "--alpha=15"],
    ["--algo=sssp", "--rmat-scale=13", "--queue-sizing=1.5"],
    ["--algo=sssp", "--rmat-scale=15", "--builtin"],
    ["--algo=pr", "--rmat-scale=14", "--max_iter=1"],
    ["--algo=pr", "--rmat-scale=14", "--max_iter=10"],
    ["--algo=pr", "--rmat-scale=14", "--max_iter=10", "--scatter"],
])
def test_frontier_driver(args):
    rc, out = _run("frontier_driver", *args)
    assert rc == 0 and "Correct." in out, out


def test_frontier_driver_on_reference_fixtures(tmp_path, golden):
    import numpy as np
    from conftest import golden_f32
    rec = golden("ref_fixture_bfs.json")
    f = _mtx(tmp_path, rec)
    for extra in ([], ["--alpha=2"], ["--builtin"]):
        # (the driver's own verdict comes from the cpu() of OUR header; the dump is compared with the labels the
        # REFERENCE's bfs_problem_t::cpu produced for this file -- tests/golden/make_golden.py)
        labels, _ = _dumped(tmp_path, "--algo=bfs", f"--file={f}", *extra)
        assert labels.tolist() == rec["bfs_labels"]
    # PR through the header enactor against what the reference's pr_enactor_t::enact left on a B200 (ref_gpu_pr_fixture.json)
    prec = golden("ref_fixture_pr.json")
    gold = golden("ref_gpu_pr_fixture.json")
    for iters in (1, 10):
        case = [c for c in gold["cases"] if c["max_iter"] == iters and not c["custom_values"]][0]
        ranks, log = _dumped(tmp_path, "--algo=pr", f"--file={_mtx(tmp_path, prec, 'p.mtx')}", f"--max_iter={iters}", dtype="float32")
        assert np.allclose(ranks, golden_f32(case, "current"), rtol=1e-4, atol=1e-6)
        for it, ln in enumerate(case["frontier_lens"]):
            assert f"finished iteration:{it} output length: {ln}" in log
    f = _mtx(tmp_path, golden("ref_fixture_sssp_directed.json"), "s.mtx")
    for extra in ([], ["--undirected"], ["--builtin"], ["--builtin", "--undirected"]):
        rc, out = _run("frontier_driver", "--algo=sssp", f"--file={f}", "--queue-sizing=2", *extra)
        assert rc == 0 and "Correct." in out, out
