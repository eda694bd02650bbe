"""CPU tests of the C-ABI graph ingest (b200_mtx_load, b200_csr_cache_*): the .mtx loader must produce exactly
what the reference's load_graph (gunrock/src/graph.hxx:96-223) produces -- checked against the UNMODIFIED reference
code compiled into oracle/_ref when it is available, and against the oracle's restatement always."""
import os

import numpy as np
import pytest

import oracle
from mini_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = ["ref_fixture_bfs.json", "ref_fixture_sssp_directed.json", "ref_fixture_sssp_undirected.json",
            "ref_fixture_pr.json"]


@pytest.fixture(scope="module", autouse=True)
def _built():
    from mini_b200 import build as b
    b.build()
    oracle.build()


def _write_mtx(path, n, edges, weights=None, comments=2, blank_every=0, long_comments=False):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        for k in range(comments):
            # (a comment longer than the reference's 100-byte fgets buffer makes load_graph exit(0) -- graph.hxx:103-111 --
            # which would end the test process: only files that are never shown to the reference get one)
            f.write(f"% comment line {k} " + ("x" * 150 if long_comments else "") + "\n")
        f.write(f"{n} {n} {len(edges)}\n")
        for k, (i, j) in enumerate(edges):
            if blank_every and k and k % blank_every == 0:
                f.write("\n")
            f.write(f"{i} {j}" + ("" if weights is None else f" {weights[k]!r}") + "\n")


def _same(got, ref: oracle.CSR):
    n, off, idx, w = got
    assert n == ref.n
    assert np.array_equal(off.astype(np.int64), np.asarray(ref.offsets, np.int64))
    assert np.array_equal(idx, np.asarray(ref.indices, np.int32))
    assert np.array_equal(w, np.asarray(ref.weights, np.float32))


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("undirected", [False, True])
def test_mtx_load_matches_reference_on_its_fixtures(tmp_path, golden, name, undirected):
    rec = golden(name)
    p = tmp_path / "g.mtx"
    with open(p, "w") as f:
        f.write(" ".join(str(x) for x in rec["mtx_header"]) + "\n")
        for e in rec["mtx_edges"]:
            f.write(" ".join(str(int(x)) if k < 2 else repr(x) for k, x in enumerate(e)) + "\n")
    got = L.load_mtx(str(p), undirected)
    _same(got, oracle.load_mtx(str(p), undirected))
    if oracle.have_ref():
        ref, _ = oracle.ref_load_graph(str(p), undirected)      # the reference's own load_graph, unmodified
        _same(got, ref)
    if undirected == rec["undirected"]:
        assert got[1].astype(np.int64).tolist() == list(rec["offsets"]) and got[2].tolist() == list(rec["indices"])


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("undirected", [False, True])
def test_mtx_load_random_graphs_without_duplicates(tmp_path, seed, undirected):
    rng = np.random.default_rng(seed)
    n = 300
    pairs = set()
    while len(pairs) < 2500:
        i, j = (int(x) for x in rng.integers(1, n + 1, 2))
        if i != j and (j, i) not in pairs:       # no duplicates after mirroring: the reference comparator stays defined
            pairs.add((i, j))
    edges = sorted(pairs, key=lambda e: rng.random())
    weights = [float(int(x)) for x in rng.integers(1, 65, len(edges))] if seed != 2 else None
    p = tmp_path / "r.mtx"
    _write_mtx(p, n + 5, edges, weights)          # 5 trailing vertices without arcs
    got = L.load_mtx(str(p), undirected)
    _same(got, oracle.load_mtx(str(p), undirected))
    assert got[0] == n + 5 and got[1][-1] == len(edges) * (2 if undirected else 1)
    if oracle.have_ref():
        ref, _ = oracle.ref_load_graph(str(p), undirected)
        _same(got, ref)


def test_mtx_load_duplicates_self_loops_and_blank_lines(tmp_path):
    """Where the reference is undefined (duplicate entries) the loader is deterministic: equal keys keep file order."""
    edges = [(2, 1), (2, 1), (1, 1), (3, 2), (2, 1), (1, 3)]
    weights = [5.0, 7.0, 1.5, 2.0, 9.0, 4.0]
    p = tmp_path / "d.mtx"
    _write_mtx(p, 3, edges, weights, blank_every=2, long_comments=True)
    n, off, idx, w = L.load_mtx(str(p), False)
    # arcs (row = j-1, col = i-1): (0,1,5) (0,1,7) (0,0,1.5) (1,2,2) (0,1,9) (2,0,4) -> sorted by (row, col), stable
    assert n == 3 and off.tolist() == [0, 4, 5, 6]
    assert idx.tolist() == [0, 1, 1, 1, 2, 0] and w.tolist() == [1.5, 5.0, 7.0, 9.0, 2.0, 4.0]


def test_mtx_load_errors(tmp_path):
    with pytest.raises(L.B200Error) as e:
        L.load_mtx(str(tmp_path / "missing.mtx"))
    assert e.value.status == 7                                    # B200_ERR_IO
    bad = tmp_path / "bad.mtx"
    for text in ("% only a comment\n", "3 3\n", "3 3 2\n1 2\n", "3 3 1\n1 9\n", "3 3 1\n0 1\n", "3 3 1\nx y\n"):
        bad.write_text(text)
        with pytest.raises(L.B200Error) as e:
            L.load_mtx(str(bad))
        assert e.value.status == 8, text                          # B200_ERR_FORMAT


def test_csr_cache_round_trip_and_corruption(tmp_path):
    g = oracle.rmat_csr(10, 8, 3, weighted=True)
    off = np.asarray(g.offsets).astype(np.uint32)
    p = str(tmp_path / "g.b200csr")
    L.write_csr_cache(p, off, g.indices, g.weights)
    n, o2, i2, w2 = L.read_csr_cache(p)
    assert n == g.n and np.array_equal(o2, off) and np.array_equal(i2, g.indices) and np.array_equal(w2, g.weights)
    L.write_csr_cache(p, off, g.indices)                          # without weights
    n, o2, i2, w2 = L.read_csr_cache(p)
    assert w2 is None and np.array_equal(i2, g.indices)
    raw = bytearray(open(p, "rb").read())
    for mutate in ("magic", "truncate", "flip", "index"):
        b = bytearray(raw)
        if mutate == "magic":
            b[0] ^= 0xFF
        elif mutate == "truncate":
            b = b[:-100]
        elif mutate == "flip":
            b[len(b) // 2] ^= 0x01                                 # payload bit flip: checksum mismatch
        else:
            b[32 + 4 * (g.n + 1)] = 0xFF                           # first column index far out of range ...
            b[32 + 4 * (g.n + 1) + 3] = 0x7F
        q = str(tmp_path / f"{mutate}.b200csr")
        open(q, "wb").write(bytes(b))
        with pytest.raises(L.B200Error) as e:
            L.read_csr_cache(q)
        assert e.value.status == 8, mutate
    with pytest.raises(L.B200Error) as e:
        L.read_csr_cache(str(tmp_path / "nope.b200csr"))
    assert e.value.status == 7


def test_empty_graph_cache(tmp_path):
    p = str(tmp_path / "e.b200csr")
    L.write_csr_cache(p, np.zeros(5, np.uint32), np.zeros(0, np.int32))
    n, off, idx, w = L.read_csr_cache(p)
    assert n == 4 and off.tolist() == [0] * 5 and len(idx) == 0 and w is None


def test_plain_c_host_uses_the_abi(tmp_path, golden):
    """examples/ingest_host.c: a C99 program (gcc, no CUDA / C++ headers) drives the C ABI -- .mtx ingest, CSR cache
    round trip, and (only with a GPU) upload + BFS.  Without a GPU it must say so instead of computing on the CPU."""
    import subprocess
    exe = tmp_path / "ingest_host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "ingest_host.c"), "-L", os.path.join(ROOT, "mini_b200"), "-lb200_frontier",
                    "-Wl,-rpath," + os.path.join(ROOT, "mini_b200"), "-o", str(exe)], check=True, capture_output=True)
    rec = golden("ref_fixture_bfs.json")
    p = tmp_path / "g.mtx"
    with open(p, "w") as f:
        f.write(" ".join(str(x) for x in rec["mtx_header"]) + "\n")
        for e in rec["mtx_edges"]:
            f.write(" ".join(str(int(x)) if k < 2 else repr(x) for k, x in enumerate(e)) + "\n")
    r = subprocess.run([str(exe), str(p), "--undirected", "--bfs", str(rec["src"])], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ref = oracle.load_mtx(str(p), True)
    assert f"n={ref.n} m={ref.m} " in r.stdout and "cache_round_trip=ok" in r.stdout
    import torch
    if torch.cuda.is_available():
        depths = oracle.bfs(ref, rec["src"])
        assert f"reached={int((depths >= 0).sum())} depth={int(depths.max())}" in r.stdout, r.stdout
    else:
        assert "bfs skipped: no CUDA device" in r.stdout
