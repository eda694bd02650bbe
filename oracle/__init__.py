"""ctypes front-end of the CPU oracle (oracle/oracle.c) and of the compiled
reference CPU code (oracle/_ref/libref_cpu.so, built from /root/reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  The product package
(mini_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libref_cpu.so")

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/libref_cpu.so when /root/reference exists)."""
    src = os.path.join(_HERE, "oracle.c")
    rnd_src, rnd_lib = os.path.join(_HERE, "stdrand.cpp"), os.path.join(_HERE, "libstdrand.so")
    stale = (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(src) or \
        (not os.path.exists(rnd_lib)) or os.path.getmtime(rnd_lib) < os.path.getmtime(rnd_src)
    gpu_bin = os.path.join(_HERE, "_ref", "ref_gpu_pr")
    need_ref = os.path.isdir("/root/reference/gunrock/src") and (
        not os.path.exists(_REF)
        or os.path.getmtime(_REF) < os.path.getmtime(os.path.join(_HERE, "ref_shim.cu"))
        or not os.path.exists(gpu_bin)
        or os.path.getmtime(gpu_bin) < os.path.getmtime(os.path.join(_HERE, "ref_gpu_driver.cu")))
    if force or stale or need_ref:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_rmat_pairs.argtypes = [C.c_int, C.c_int64, C.c_uint64, _i32p, _i32p]
        L.orc_rmat_pairs.restype = None
        L.orc_build_csr.argtypes = [C.c_int64, C.c_int64, _i32p, _i32p, C.c_int, _i64p, _i32p,
                                    C.c_void_p, C.c_uint64]
        L.orc_build_csr.restype = None
        L.orc_bfs.argtypes = [C.c_int64, _i64p, _i32p, C.c_int32, _i32p]
        L.orc_bfs.restype = None
        L.orc_sssp_ref_preds.argtypes = [C.c_int64, _i64p, _i32p, _f32p, C.c_int32, _i32p, C.c_void_p]
        L.orc_sssp_ref_preds.restype = None
        L.orc_sssp_dist.argtypes = [C.c_int64, _i64p, _i32p, _f32p, C.c_int32, _f32p]
        L.orc_sssp_dist.restype = None
        L.orc_sssp_dist_f32.argtypes = [C.c_int64, _i64p, _i32p, _f32p, C.c_int32, _f32p]
        L.orc_sssp_dist_f32.restype = None
        L.orc_neighborhood_reduce_f64.argtypes = [C.c_int64, _i32p, _i64p, _i32p, _f64p, C.c_int,
                                                  C.c_double, _f64p, C.c_void_p]
        L.orc_neighborhood_reduce_f64.restype = None
        L.orc_pr.argtypes = [C.c_int64, _i64p, _i32p, C.c_int, C.c_int, _f32p, _f32p, _i64p, C.c_void_p]
        L.orc_pr.restype = C.c_int
        L.orc_bfs_push_level.argtypes = [_i64p, _i32p, _i32p, C.c_int64, C.c_int32, _i32p, _i32p]
        L.orc_bfs_push_level.restype = C.c_int64
        L.orc_reached_arcs_i32.argtypes = [C.c_int64, _i64p, _i32p]
        L.orc_reached_arcs_i32.restype = C.c_int64
        L.orc_kcore.argtypes = [C.c_int64, _i64p, _i32p, _i32p]
        L.orc_kcore.restype = C.c_int32
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def have_ref() -> bool:
    build()
    return os.path.exists(_REF)


def ref():
    """The unmodified reference CPU code (None if it was never built)."""
    global _ref
    if _ref is None:
        if not have_ref():
            return None
        R = C.CDLL(_REF)
        R.ref_bfs_cpu.argtypes = [C.c_int, C.c_int, _i32p, _i32p, C.c_int, _i32p]
        R.ref_bfs_cpu.restype = C.c_double
        R.ref_sssp_cpu.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f32p, C.c_int, _i32p]
        R.ref_sssp_cpu.restype = C.c_double
        R.ref_load_graph.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        R.ref_load_graph.restype = C.c_int
        _ref = R
    return _ref


def have_ref_gpu() -> bool:
    """The unmodified reference GPU primitives (oracle/_ref/ref_gpu_{bfs,sssp,pr}, ref_gpu_driver.cu)."""
    build()
    return all(os.path.exists(os.path.join(_HERE, "_ref", f"ref_gpu_{p}")) for p in ("bfs", "sssp", "pr"))


def ref_gpu(prim: str, g: "CSR", src: int = 0, alpha: float | None = None, max_iter: int = 10, runs: int = 1,
            queue_sizing: float = 1.0, values=None, timeout: float = 600.0):
    """Runs the reference's own GPU enactor (bfs: enact_pushpull, sssp: enact, pr: enact) on the CSR in a child
    process (the reference exit()s on frontier overflow and prints from its enactors).  Needs a GPU.
    Returns (dict of output arrays, list of wall-clock seconds per run as the reference's tests time them)."""
    import json
    import tempfile
    assert g.m < 2 ** 31 and g.n < 2 ** 31, "the reference's CSR is int32"
    exe = os.path.join(_HERE, "_ref", f"ref_gpu_{prim}")
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(np.array([g.n, g.m], np.int64).tobytes())
            f.write(np.array([0 if g.weights is None else 1, 1], np.int32).tobytes())
            f.write(g.offsets.astype(np.int32).tobytes())
            f.write(g.indices.tobytes())
            if g.weights is not None:
                f.write(g.weights.tobytes())
        cmd = [exe, prim, fin, fout, f"--src={src}", f"--max_iter={max_iter}", f"--runs={runs}",
               f"--queue-sizing={queue_sizing}"]
        if alpha is not None:
            cmd.append(f"--alpha={alpha!r}")
        if values is not None:
            fv = os.path.join(td, "values.bin")
            np.ascontiguousarray(values, np.float32).tofile(fv)
            cmd.append(f"--values={fv}")
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not last:
            raise RuntimeError(f"ref_gpu_{prim} failed rc={r.returncode}: {r.stdout[-400:]} {r.stderr[-400:]}")
        info = json.loads(last[-1])
        raw = np.fromfile(fout, np.uint8)
    n = g.n
    if prim == "bfs":
        out = {"labels": raw[:4 * n].view(np.int32).copy()}
    elif prim == "sssp":
        out = {"labels": raw[:4 * n].view(np.float32).copy(), "preds": raw[4 * n:8 * n].view(np.int32).copy()}
    else:
        out = {"current": raw[:4 * n].view(np.float32).copy(), "reduced": raw[4 * n:8 * n].view(np.float32).copy()}
    out["stdout"] = r.stdout
    return out, info["elapsed_s"]


# --------------------------------------------------------------------------- graphs
class CSR:
    """Host CSR: int64 offsets[n+1], int32 indices[m], optional float32 weights[m]."""

    def __init__(self, n, offsets, indices, weights=None):
        self.n = int(n)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.weights = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        self.m = int(self.offsets[-1])
        assert self.indices.shape[0] == self.m


def rmat_pairs(scale: int, edge_factor: int = 16, seed: int = 1):
    npairs = edge_factor << scale
    src = np.empty(npairs, np.int32)
    dst = np.empty(npairs, np.int32)
    lib().orc_rmat_pairs(scale, npairs, seed, src, dst)
    return src, dst


def build_csr(n: int, src, dst, symmetrize: bool = True, weighted: bool = False, wseed: int = 7) -> CSR:
    src = np.ascontiguousarray(src, np.int32)
    dst = np.ascontiguousarray(dst, np.int32)
    m = src.shape[0] * (2 if symmetrize else 1)
    offsets = np.empty(n + 1, np.int64)
    indices = np.empty(m, np.int32)
    weights = np.empty(m, np.float32) if weighted else None
    lib().orc_build_csr(n, src.shape[0], src, dst, int(symmetrize), offsets, indices,
                        None if weights is None else weights.ctypes.data, wseed)
    return CSR(n, offsets, indices, weights)


def rmat_csr(scale: int, edge_factor: int = 16, seed: int = 1, weighted: bool = False, wseed: int = 7) -> CSR:
    """Symmetrised RMAT CSR of SURVEY.md §8d (duplicates and self loops kept, m = 2*ef*2^scale)."""
    s, d = rmat_pairs(scale, edge_factor, seed)
    return build_csr(1 << scale, s, d, True, weighted, wseed)


def load_mtx(path: str, undirected: bool) -> CSR:
    """Restates gunrock/src/graph.hxx:96-223 (load_graph): a line ``i j [w]`` becomes arc
    j-1 -> i-1 (CSR rows are built over mtx column 2, :139-172), weight 1.0 if absent
    (:120-127, _random_edge_value=false), ``undirected`` appends every reverse (:130-137),
    arcs sorted by (row, col).  Valid for files without duplicate edges (the reference's
    comparator is UB otherwise)."""
    with open(path) as f:
        lines = [ln for ln in f.read().splitlines() if ln.strip()]
    k = 0
    while lines[k].startswith("%"):
        k += 1
    h, _w, ne = (int(x) for x in lines[k].split()[:3])
    a, b, w = [], [], []
    for ln in lines[k + 1:k + 1 + ne]:
        t = ln.split()
        a.append(int(t[0]) - 1)
        b.append(int(t[1]) - 1)
        w.append(float(t[2]) if len(t) > 2 else 1.0)
    a, b, w = np.array(a, np.int64), np.array(b, np.int64), np.array(w, np.float32)
    if undirected:
        a, b, w = np.concatenate([a, b]), np.concatenate([b, a]), np.concatenate([w, w])
    order = np.lexsort((a, b))          # primary key: column 2 (b), secondary: column 1 (a)
    rows, cols, w = b[order], a[order], w[order]
    offsets = np.zeros(h + 1, np.int64)
    np.add.at(offsets, rows + 1, 1)
    offsets = np.cumsum(offsets)
    return CSR(h, offsets, cols.astype(np.int32), w)


# --------------------------------------------------------------------------- algorithms
def bfs(g: CSR, src: int = 0) -> np.ndarray:
    labels = np.full(g.n, -1, np.int32)
    lib().orc_bfs(g.n, g.offsets, g.indices, src, labels)
    return labels


def bfs_timed(g: CSR, src: int = 0):
    labels = np.full(g.n, -1, np.int32)
    t0 = time.perf_counter()
    lib().orc_bfs(g.n, g.offsets, g.indices, src, labels)
    return labels, time.perf_counter() - t0


def sssp_dist(g: CSR, src: int = 0) -> np.ndarray:
    out = np.empty(g.n, np.float32)
    lib().orc_sssp_dist(g.n, g.offsets, g.indices, g.weights, src, out)
    return out


def sssp_dist_f32(g: CSR, src: int = 0) -> np.ndarray:
    """orc_sssp_dist for weights that are not integers: Dijkstra in fp32 under the functor's rounding."""
    out = np.empty(g.n, np.float32)
    lib().orc_sssp_dist_f32(g.n, g.offsets, g.indices, g.weights, src, out)
    return out


def sssp_ref_preds(g: CSR, src: int = 0):
    preds = np.full(g.n, -1, np.int32)
    dist = np.empty(g.n, np.int32)
    lib().orc_sssp_ref_preds(g.n, g.offsets, g.indices, g.weights, src, preds, dist.ctypes.data)
    return preds, dist


def kcore(g: CSR):
    """(num_cores, largest k-core) of the reference's peel, kcore_problem.hxx:54-105 restated (oracle.c orc_kcore).
    PARITY UNPINNED: the reference's kcore_problem.hxx does not compile (kcore_problem.hxx:44), so its cpu() could
    not be built here to check this restatement; the hand-computed cases in tests/test_oracle.py stand in."""
    cores = np.empty(g.n, np.int32)
    largest = lib().orc_kcore(g.n, g.offsets, g.indices, cores)
    return cores, int(largest)


_stdrand = None


def mt19937_hashes(n: int, prime: int, restart: bool = False):
    """The next n draws of std::uniform_int_distribution<int>(0, prime) from ONE default-seeded std::mt19937 -- what
    mgpu::fill_random(0, prime, n, false, ctx) hands the colouring problem (memory.hxx:112-129,
    coloring_problem.hxx:44,50).  Not a restatement: oracle/stdrand.cpp calls the installed libstdc++ (whose
    distribution algorithm differs between releases).  restart = True re-seeds the shared engine first (a new process
    in the reference's terms)."""
    global _stdrand
    if _stdrand is None:
        build()
        _stdrand = C.CDLL(os.path.join(_HERE, "libstdrand.so"))
        _stdrand.orc_stdrand_uniform.argtypes = [C.c_longlong, C.c_int, C.c_int, _i32p]
        _stdrand.orc_stdrand_uniform.restype = None
        _stdrand.orc_stdrand_reset.restype = None
    if restart:
        _stdrand.orc_stdrand_reset()
    out = np.empty(n, np.int32)
    _stdrand.orc_stdrand_uniform(n, 0, prime, out)
    return out


def coloring(g: CSR, prime: int = 15485863, max_iter: int = 10):
    """Hash-extrema colouring of coloring_enactor.hxx:43-92 + coloring_functor.hxx:10-70, restated with numpy:
    every iteration the uncoloured vertices whose hash is below (above) every UNCOLOURED neighbour's take colour
    2*it+1 (2*it+2); all hashes are redrawn.  Returns (colors, uncoloured count after every iteration)."""
    n = g.n
    deg = np.diff(g.offsets)
    rows = np.repeat(np.arange(n, dtype=np.int64), deg)
    colors = np.zeros(n, np.int32)
    frontier = np.arange(n, dtype=np.int64)
    hashes = mt19937_hashes(n, prime, restart=True)
    lens = []
    it = 0
    imin, imax = np.iinfo(np.int32).min, np.iinfo(np.int32).max
    while len(frontier) and it < max_iter:
        live = colors[g.indices] == 0
        hv = hashes[g.indices].astype(np.int64)
        rmax = np.full(n, imin, np.int64)
        rmin = np.full(n, imax, np.int64)
        np.maximum.at(rmax, rows, np.where(live, hv, imin))
        np.minimum.at(rmin, rows, np.where(live, hv, imax))
        h = hashes[frontier].astype(np.int64)
        col = np.where(h < rmin[frontier], 2 * it + 1, np.where(h > rmax[frontier], 2 * it + 2, 0))
        colors[frontier[col > 0]] = col[col > 0]
        frontier = frontier[col == 0]
        lens.append(len(frontier))
        it += 1
        hashes = mt19937_hashes(n, prime)
    return colors, lens


def neighborhood_reduce(g: CSR, frontier, values, op: str = "plus", identity: float = 0.0):
    frontier = np.ascontiguousarray(frontier, np.int32)
    values = np.ascontiguousarray(values, np.float64)
    red = np.empty(frontier.shape[0], np.float64)
    asum = np.empty(frontier.shape[0], np.float64)
    lib().orc_neighborhood_reduce_f64(frontier.shape[0], frontier, g.offsets, g.indices, values,
                                      {"plus": 0, "min": 1, "max": 2}[op], identity, red,
                                      asum.ctypes.data)
    return red, asum


def pr(g: CSR, max_iter: int = 10, scatter: bool = False, init=None):
    cur = np.empty(g.n, np.float32)
    red = np.empty(g.n, np.float32)
    lens = np.zeros(max(max_iter, 1), np.int64)
    init = None if init is None else np.ascontiguousarray(init, np.float32)
    it = lib().orc_pr(g.n, g.offsets, g.indices, max_iter, int(scatter), cur, red, lens,
                      None if init is None else init.ctypes.data)
    return cur, red, lens[:it]


def bfs_push_level(g: CSR, frontier, iteration: int, labels: np.ndarray):
    frontier = np.ascontiguousarray(frontier, np.int32)
    nxt = np.empty(g.n, np.int32)
    k = lib().orc_bfs_push_level(g.offsets, g.indices, frontier, frontier.shape[0], iteration, labels, nxt)
    return nxt[:k].copy()


def reached_arcs(g: CSR, labels) -> int:
    return int(lib().orc_reached_arcs_i32(g.n, g.offsets, np.ascontiguousarray(labels, np.int32)))


def num_threads() -> int:
    return int(lib().orc_num_threads())


# --------------------------------------------------------------------------- reference wrappers
def ref_bfs(g: CSR, src: int = 0):
    """Unmodified bfs_problem_t::cpu (int32 CSR only). Returns (labels, seconds)."""
    R = ref()
    assert R is not None and g.m < 2 ** 31
    labels = np.empty(g.n, np.int32)
    t = R.ref_bfs_cpu(g.n, g.m, g.offsets.astype(np.int32), g.indices, src, labels)
    return labels, t


def ref_sssp_preds(g: CSR, src: int = 0):
    R = ref()
    assert R is not None and g.m < 2 ** 31
    preds = np.empty(g.n, np.int32)
    t = R.ref_sssp_cpu(g.n, g.m, g.offsets.astype(np.int32), g.indices, g.weights, src, preds)
    return preds, t


def ref_load_graph(path: str, undirected: bool):
    R = ref()
    assert R is not None
    n, m, eq = C.c_int(), C.c_int(), C.c_int()
    rc = R.ref_load_graph(path.encode(), int(undirected), C.byref(n), C.byref(m), None, None, None, None)
    assert rc == 0, path
    off = np.empty(n.value + 1, np.int32)
    ind = np.empty(m.value, np.int32)
    w = np.empty(m.value, np.float32)
    R.ref_load_graph(path.encode(), int(undirected), C.byref(n), C.byref(m),
                     off.ctypes.data, ind.ctypes.data, w.ctypes.data, C.byref(eq))
    return CSR(n.value, off, ind, w), bool(eq.value)
