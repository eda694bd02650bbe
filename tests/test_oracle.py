"""Pins the CPU oracle (oracle/oracle.c) against
 (a) the committed golden vectors produced by the reference's own CPU code
     (tests/golden/make_golden.py), and
 (b) when oracle/_ref/libref_cpu.so is present, the reference code itself, live.
Reference: gunrock/src/bfs/bfs_problem.hxx:52-72, gunrock/src/sssp/sssp_problem.hxx:59-88,
gunrock/src/graph.hxx:96-223, gunrock/tests/{bfs,sssp}/test_*.cu:44-52."""
import hashlib
import os

import numpy as np
import pytest

import oracle

FIXTURES = ["ref_fixture_bfs.json", "ref_fixture_sssp_directed.json",
            "ref_fixture_sssp_undirected.json", "ref_fixture_pr.json"]
RMATS = ["ref_rmat_s8.json", "ref_rmat_s10.json", "ref_rmat_s16.json", "ref_rmat_s12_seed3.json"]


def _write_mtx(tmp_path, rec):
    p = tmp_path / "g.mtx"
    with open(p, "w") as f:
        f.write(" ".join(str(x) for x in rec["mtx_header"]) + "\n")
        for e in rec["mtx_edges"]:
            f.write(" ".join(str(int(x)) if k < 2 else repr(x) for k, x in enumerate(e)) + "\n")
    return str(p)


def _csr_from(rec):
    return oracle.CSR(rec["n"], rec["offsets"], rec["indices"], rec["weights"])


@pytest.mark.parametrize("name", FIXTURES)
def test_load_mtx_matches_reference_load_graph(name, golden, tmp_path):
    rec = golden(name)
    g = oracle.load_mtx(_write_mtx(tmp_path, rec), rec["undirected"])
    assert g.n == rec["n"] and g.m == rec["m"]
    assert g.offsets.tolist() == rec["offsets"]
    assert g.indices.tolist() == rec["indices"]
    assert g.weights.tolist() == rec["weights"]
    assert rec["csc_equals_csr"]          # device CSC is always a copy of CSR


def test_survey_known_answers(golden):
    """The §4 table of SURVEY.md (produced with the reference's own code)."""
    r = golden("ref_fixture_bfs.json")
    assert r["offsets"] == [0, 4, 8, 14, 18, 23, 27, 30] and r["bfs_labels"] == [0, 1, 1, 1, 2, 2, 2]
    r = golden("ref_fixture_sssp_directed.json")
    assert r["indices"] == [1, 2, 3, 2, 4, 3, 4, 5, 5, 6, 5, 6, 6]
    assert r["weights"] == [3, 1, 3, 2, 7, 1, 5, 2, 4, 3, 1, 2, 1]
    assert r["sssp_preds"] == [-1, 0, 0, 2, 2, 2, 5]
    assert golden("ref_fixture_sssp_undirected.json")["sssp_preds"] == [-1, 0, 0, 2, 5, 2, 5]


@pytest.mark.parametrize("name", FIXTURES)
def test_bfs_and_sssp_on_reference_fixtures(name, golden):
    rec = golden(name)
    g = _csr_from(rec)
    assert oracle.bfs(g, rec["src"]).tolist() == rec["bfs_labels"]
    preds, dist_i = oracle.sssp_ref_preds(g, rec["src"])
    assert preds.tolist() == rec["sssp_preds"]
    # distances: Dijkstra under the GPU functor's rule == the reference cpu()'s int distances
    d = oracle.sssp_dist(g, rec["src"])
    reach = dist_i != np.iinfo(np.int32).max
    assert np.array_equal(d[reach], dist_i[reach].astype(np.float32))
    assert np.all(d[~reach] == np.finfo(np.float32).max)


def test_sssp_distances_known(golden):
    g = _csr_from(golden("ref_fixture_sssp_directed.json"))
    assert oracle.sssp_dist(g, 0).tolist() == [0, 3, 1, 2, 6, 3, 4]
    g = _csr_from(golden("ref_fixture_sssp_undirected.json"))
    assert oracle.sssp_dist(g, 0).tolist() == [0, 3, 1, 2, 4, 3, 4]


@pytest.mark.parametrize("name", RMATS)
def test_rmat_generator_and_traversals_match_golden(name, golden):
    rec = golden(name)
    g = oracle.rmat_csr(rec["scale"], rec["edge_factor"], rec["seed"], weighted=True, wseed=rec["wseed"])
    assert g.n == rec["n"] and g.m == rec["m"] == 2 * rec["edge_factor"] << rec["scale"]
    sha = hashlib.sha256(g.offsets.tobytes() + g.indices.tobytes() + g.weights.tobytes()).hexdigest()
    assert sha == rec["csr_sha256"]
    labels = oracle.bfs(g, rec["src"])
    assert hashlib.sha256(labels.tobytes()).hexdigest() == rec["bfs_labels_sha256"]
    assert np.bincount(labels + 1).tolist() == rec["bfs_labels_hist"]
    preds, _ = oracle.sssp_ref_preds(g, rec["src"])
    assert hashlib.sha256(preds.tobytes()).hexdigest() == rec["sssp_preds_sha256"]
    if rec["bfs_labels"] is not None:
        assert labels.tolist() == rec["bfs_labels"] and preds.tolist() == rec["sssp_preds"]


def test_rmat_csr_structure():
    g = oracle.rmat_csr(9, 16, 1, weighted=True)
    n, m = g.n, g.m
    assert g.offsets[0] == 0 and g.offsets[-1] == m and np.all(np.diff(g.offsets) >= 0)
    rows = np.repeat(np.arange(n), np.diff(g.offsets))
    key = rows.astype(np.int64) << 32 | g.indices
    assert np.all(np.diff(key) >= 0)                              # sorted by (src, dst)
    rkey = np.sort(g.indices.astype(np.int64) << 32 | rows)
    assert np.array_equal(np.sort(key), rkey)                     # symmetric multiset
    assert g.weights.min() >= 1 and g.weights.max() <= 64 and np.all(g.weights == np.floor(g.weights))
    # same weight both directions
    w_fwd = dict(zip(key.tolist(), g.weights.tolist()))
    for k in list(w_fwd)[:2000]:
        assert w_fwd[k] == w_fwd[(k & 0xFFFFFFFF) << 32 | k >> 32]


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_against_compiled_reference(seed):
    g = oracle.rmat_csr(11, 8, seed, weighted=True)
    for src in (0, 1, 17):
        ref_labels, _ = oracle.ref_bfs(g, src)
        assert np.array_equal(oracle.bfs(g, src), ref_labels)
        ref_preds, _ = oracle.ref_sssp_preds(g, src)
        assert np.array_equal(oracle.sssp_ref_preds(g, src)[0], ref_preds)


def test_neighborhood_reduce_semantics():
    """mgpu tests/test_segreduce.cu:40-58: empty segment => init, init not folded into non-empty."""
    g = oracle.CSR(4, [0, 2, 2, 5, 6], [1, 2, 0, 1, 3, 3])
    vals = np.array([1.0, 10.0, 100.0, 1000.0])
    red, asum = oracle.neighborhood_reduce(g, [0, 1, 2, 3, 1], vals, "plus", identity=-1.0)
    assert red.tolist() == [110.0, -1.0, 1011.0, 1000.0, -1.0]
    red, _ = oracle.neighborhood_reduce(g, [2, 0], vals, "min", identity=7.0)
    assert red.tolist() == [1.0, 10.0]
    red, _ = oracle.neighborhood_reduce(g, [2, 1], vals, "max", identity=7.0)
    assert red.tolist() == [1000.0, 7.0]


def test_pr_driver_on_fixture(golden):
    """pr_enactor.hxx:41-77 restated; iteration 0 (frontier = iota) is mode independent."""
    g = _csr_from(golden("ref_fixture_pr.json"))
    cur_a, red_a, lens_a = oracle.pr(g, 1, scatter=False)
    cur_b, red_b, lens_b = oracle.pr(g, 1, scatter=True)
    assert np.array_equal(cur_a, cur_b) and np.array_equal(red_a, red_b)
    deg = np.diff(g.offsets).astype(np.float32)
    assert np.allclose(red_a, 0.15 * deg)
    assert np.allclose(cur_a, 0.15 + 0.85 * 0.15)
    cur, red, lens = oracle.pr(g, 50, scatter=True)
    assert len(lens) <= 50 and np.all(np.isfinite(cur))


# ---------------------------------------------------------------- pinned by the reference's own GPU code
# tests/golden/ref_gpu_*.json: outputs of the UNMODIFIED reference GPU primitives compiled for sm_100 and run on a
# B200 (oracle/ref_gpu_driver.cu, tests/golden/make_golden_gpu.py).  They pin the parts of the oracle the
# reference's tests never validate: neighbourhood sums, the PR driver (incl. its slot-indexed read)
# and the SSSP distances.
REF_GPU_PR = [("ref_gpu_pr_fixture.json", "fixture"), ("ref_gpu_pr_rmat_s12.json", "rmat12")]


def _pr_graph(golden, which):
    if which == "fixture":
        return _csr_from(golden("ref_fixture_pr.json"))
    return oracle.rmat_csr(12, 16, 1)


@pytest.mark.parametrize("name,which", REF_GPU_PR)
def test_neighborhood_reduce_matches_reference_gpu(golden, name, which):
    """One enact() iteration from non-uniform ranks: d_reduced_ranks = the neighbourhood sums of neighborhood.hxx:47-58
    computed by the reference's lbs_segreduce.  Tolerance: 1e-5 * sum|terms| + 1e-6 per slot."""
    from conftest import golden_custom_values, golden_f32
    rec = golden(name)
    g = _pr_graph(golden, which)
    case = [c for c in rec["cases"] if c["max_iter"] == 1 and c["custom_values"]][0]
    vals = golden_custom_values(g.n)
    red, asum = oracle.neighborhood_reduce(g, np.arange(g.n, dtype=np.int32), vals.astype(np.float64))
    ref = golden_f32(case, "reduced").astype(np.float64)
    assert np.all(np.abs(red - ref) <= 1e-5 * asum + 1e-6)
    # and the filter's update of the ranks (pr_functor.hxx:11-17) from those sums
    deg = np.diff(g.offsets).astype(np.float32)
    new = np.where(deg > 0, np.float32(0.15) + np.float32(0.85) * ref.astype(np.float32) / np.maximum(deg, 1), np.float32(0.15))
    assert np.allclose(golden_f32(case, "current"), new, rtol=1e-6, atol=0)
    assert case["frontier_lens"] == [int((np.abs(new - vals) > np.float32(0.001) * vals).sum())]


@pytest.mark.parametrize("name,which", REF_GPU_PR)
def test_pr_driver_matches_reference_gpu(golden, name, which):
    """pr_enactor_t::enact for 1 / 3 / 10 iterations, default and non-uniform initial ranks: identical frontier
    length after every iteration, ranks and sums within rel 1e-4 of the reference GPU run."""
    from conftest import golden_custom_values, golden_f32
    rec = golden(name)
    g = _pr_graph(golden, which)
    for case in rec["cases"]:
        init = golden_custom_values(g.n) if case["custom_values"] else None
        cur, red, lens = oracle.pr(g, case["max_iter"], scatter=False, init=init)
        assert lens.tolist() == case["frontier_lens"], (case["max_iter"], case["custom_values"])
        assert np.allclose(cur, golden_f32(case, "current"), rtol=1e-4, atol=1e-6)
        assert np.allclose(red, golden_f32(case, "reduced"), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name,which", [("ref_gpu_traversal_fixture_bfs.json", "ref_fixture_bfs.json"),
                                        ("ref_gpu_traversal_fixture_sssp.json", "ref_fixture_sssp_undirected.json"),
                                        ("ref_gpu_traversal_rmat_s12.json", "rmat12")])
def test_traversals_match_reference_gpu(golden, name, which):
    """BFS depths (push-only and with the push -> pull switch forced) and SSSP DISTANCES (d_labels, which the
    reference's own test never compares) of the reference GPU enactors, bit-exact."""
    import hashlib
    rec = golden(name)
    g = oracle.rmat_csr(12, 16, 1, weighted=True) if which == "rmat12" else _csr_from(golden(which))
    for c in rec["cases"]:
        lab = oracle.bfs(g, c["src"])
        assert hashlib.sha256(lab.tobytes()).hexdigest() == c["bfs_labels_sha256"] == c["bfs_pushpull_labels_sha256"]
        assert hashlib.sha256(oracle.sssp_dist(g, c["src"]).tobytes()).hexdigest() == c["sssp_dist_sha256"]


# ---------------------------------------------------------------- kcore / coloring, oracle side
def test_kcore_restatement_hand_cases():
    """kcore_problem.hxx:54-105 restated (unpinned: the reference's kcore headers do not compile, see oracle.kcore).
    Hand-computed peels, including the reference's quirk: a vertex whose degree reaches 0 only because its neighbours
    were peeled is never numbered."""
    star = oracle.CSR(4, [0, 3, 4, 5, 6], [1, 2, 3, 0, 0, 0])                     # centre 0, three leaves
    cores, largest = oracle.kcore(star)
    assert cores.tolist() == [0, 1, 1, 1] and largest == 1                        # the centre keeps 0 (quirk)
    tri = oracle.build_csr(4, np.array([0, 1, 2, 0], np.int32), np.array([1, 2, 0, 3], np.int32), True)
    cores, largest = oracle.kcore(tri)                                            # triangle 0-1-2 + pendant 3
    assert cores.tolist() == [2, 2, 2, 1] and largest == 2
    k5 = oracle.build_csr(6, *np.array([(a, b) for a in range(5) for b in range(a)], np.int32).T.copy(), True)
    cores, largest = oracle.kcore(k5)                                             # K5 + an isolated vertex
    assert cores.tolist() == [4, 4, 4, 4, 4, 0] and largest == 4
    path = oracle.build_csr(5, np.arange(4, dtype=np.int32), np.arange(1, 5, dtype=np.int32), True)
    cores, largest = oracle.kcore(path)      # k = 2 peels the two ends (core 1); the next round peels 1 and 3; vertex 2 drops to 0 un-numbered
    assert cores.tolist() == [1, 1, 0, 1, 1] and largest == 1


def test_kcore_on_reference_fixture(golden):
    rec = golden("ref_fixture_kcore.json")
    cores, largest = oracle.kcore(_csr_from(rec))
    # tests/kcore/test_kcore.mtx lists every edge in both directions and load_graph(undirected) doubles it again:
    # a multigraph whose degrees are twice the simple graph's
    assert cores.tolist() == [4, 4, 6, 6, 6, 4, 6, 6, 4] and largest == 6


def test_mt19937_hash_stream():
    """The hash stream of coloring_problem_t (mgpu::fill_random, memory.hxx:112-129) comes from the installed C++
    library itself (oracle/stdrand.cpp): default-seeded std::mt19937 (first raw words 3499211612, 581869302, ...) through
    std::uniform_int_distribution<int>(0, prime); the engine is carried from call to call, not re-seeded."""
    full = oracle.mt19937_hashes(4, 2 ** 31 - 1, restart=True)        # the whole non-negative int range
    assert full.min() >= 0
    a = oracle.mt19937_hashes(6, 15485863, restart=True)
    b = oracle.mt19937_hashes(6, 15485863)
    c = oracle.mt19937_hashes(12, 15485863, restart=True)
    assert a.min() >= 0 and a.max() <= 15485863
    assert np.array_equal(np.concatenate([a, b]), c)                  # one engine, carried on


@pytest.mark.parametrize("which", ["fixture", "rmat10"])
def test_coloring_restatement_is_a_proper_partial_colouring(golden, which):
    g = _csr_from(golden("ref_fixture_coloring.json")) if which == "fixture" else oracle.rmat_csr(10, 16, 1)
    colors, lens = oracle.coloring(g, 15485863, 10)
    assert lens == sorted(lens, reverse=True) and lens[0] < g.n
    rows = np.repeat(np.arange(g.n), np.diff(g.offsets))
    cols = g.indices
    both = (colors[rows] > 0) & (colors[cols] > 0) & (rows != cols)
    assert not np.any(colors[rows][both] == colors[cols][both])    # two coloured end points never share a colour
    assert int((colors == 0).sum()) == (lens[-1] if lens else g.n)
    self_loop = np.zeros(g.n, bool)
    self_loop[rows[rows == cols]] = True
    assert np.all(colors[self_loop] == 0)    # a vertex with a self loop is its own neighbour: never a strict extremum


def test_push_level_restates_bfs():
    g = oracle.rmat_csr(10, 16, 1)
    labels = np.full(g.n, -1, np.int32)
    labels[0] = 0
    f = np.array([0], np.int32)
    it = 0
    while len(f):
        f = oracle.bfs_push_level(g, f, it, labels)
        it += 1
    assert np.array_equal(labels, oracle.bfs(g, 0))


def test_sssp_fp32_oracle_is_the_fixed_point_of_the_relax_rule():
    """orc_sssp_dist_f32 (Dijkstra under fl(dist[src] + w), sssp_functor.hxx:20-29) equals the integer oracle on integer
    weights and, on fractional weights, the fixed point of synchronous fp32 relaxation sweeps (what any order of the
    GPU enactor's atomicMin relaxations converges to)."""
    rng = np.random.default_rng(3)
    n = 3000
    s = rng.integers(0, n, 9000).astype(np.int32)
    d = rng.integers(0, n, 9000).astype(np.int32)
    g = oracle.build_csr(n, s, d, True, True)
    gf = oracle.CSR(n, g.offsets, g.indices, (rng.random(g.m, dtype=np.float32) * 3.0 + 0.01).astype(np.float32))
    big = np.finfo(np.float32).max
    src_of = np.repeat(np.arange(n), np.diff(g.offsets))
    for src in (0, 2999):
        assert oracle.sssp_dist_f32(g, src).tobytes() == oracle.sssp_dist(g, src).tobytes()
        want = oracle.sssp_dist_f32(gf, src)
        dist = np.full(n, big, np.float32)
        dist[src] = 0
        while True:
            nd = dist[src_of] + gf.weights
            nd[dist[src_of] == big] = big
            old = dist.copy()
            np.minimum.at(dist, gf.indices, nd)
            if np.array_equal(old, dist):
                break
        assert dist.tobytes() == want.tobytes()
