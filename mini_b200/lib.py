"""ctypes binding of include/b200_frontier.h (one function per C entry point)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("B200_FRONTIER_LIB") or os.path.join(_HERE, "libb200_frontier.so")   # env: tuning builds

BFS_PUSH, BFS_REF_ALPHA, BFS_BEAMER = 0, 1, 2
ADV_IDEMPOTENT, ADV_NO_OUTPUT, ADV_RAW_OUTPUT = 1, 2, 4
OP_PLUS, OP_MIN, OP_MAX = 0, 1, 2
PROBLEM_BFS, PROBLEM_SSSP, PROBLEM_PR = 1, 2, 3
MAX_LEVELS = 512
ADVANCE_QUAD, ADVANCE_LBS, ADVANCE_QUAD_RESCAN = 0, 1, 2
LOOP_GRAPH, LOOP_HOST = 0, 1


class B200Error(RuntimeError):
    def __init__(self, status: int, what: str, cuda_error: int = 0):
        super().__init__(f"{what}: status {status} ({_status_string(status)})"
                         + (f", cudaError {cuda_error}" if cuda_error else ""))
        self.status = status
        self.cuda_error = cuda_error


class CGraph(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("row_offsets", C.c_void_p), ("col_indices", C.c_void_p),
                ("col_values", C.c_void_p), ("col_offsets", C.c_void_p), ("row_indices", C.c_void_p),
                ("row_values", C.c_void_p), ("no_in_arc_bitmap", C.c_void_p), ("first_in_neighbor", C.c_void_p),
                ("hot_ids", C.c_void_p), ("hot_indices", C.c_void_p), ("hot_count", C.c_int64)]


class CHostCSR(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("row_offsets", C.c_void_p), ("col_indices", C.c_void_p),
                ("col_values", C.c_void_p)]


class CProblem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("labels", C.c_void_p), ("preds", C.c_void_p),
                ("dist", C.c_void_p), ("weights", C.c_void_p), ("visited", C.c_void_p),
                ("current_ranks", C.c_void_p), ("reduced_ranks", C.c_void_p), ("degrees", C.c_void_p),
                ("visited_bitmap", C.c_void_p)]


class CLevelStat(C.Structure):
    _fields_ = [("direction", C.c_int32), ("reserved", C.c_int32), ("frontier_len", C.c_int64), ("arcs", C.c_int64),
                ("discovered", C.c_int64), ("advance_ms", C.c_float), ("level_ms", C.c_float)]


class CStats(C.Structure):
    _fields_ = [("collect_timing", C.c_int32), ("num_levels", C.c_int32), ("reached", C.c_int64),
                ("total_arcs", C.c_int64), ("launches", C.c_int64), ("device_ms", C.c_float),
                ("level_loop", C.c_int32), ("level", CLevelStat * MAX_LEVELS)]


_lib = None


def library_path() -> str:
    return _LIB_PATH


def _status_string(status: int) -> str:
    if _lib is None:
        return "?"
    return _lib.b200_status_string(status).decode()


def load_library():
    """Loads libb200_frontier.so.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(mini_b200 has no CPU or PyTorch fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
    pi64, pi32 = C.POINTER(C.c_int64), C.POINTER(C.c_int)
    pg, pp, ps = C.POINTER(CGraph), C.POINTER(CProblem), C.POINTER(CStats)
    sig = {
        "b200_abi_version": ([], i32),
        "b200_status_string": ([i32], C.c_char_p),
        "b200_last_cuda_error": ([], i32),
        "b200_device_count": ([pi32], i32),
        "b200_ctx_create": ([C.POINTER(vp), i32, vp], i32),
        "b200_ctx_destroy": ([vp], i32),
        "b200_ctx_reserve": ([vp, i64], i32),
        "b200_ctx_sync": ([vp], i32),
        "b200_ctx_num_sms": ([vp, pi32], i32),
        "b200_ctx_l2_pin": ([vp, vp, i64], i32),
        "b200_ctx_set_advance_impl": ([vp, i32], i32),
        "b200_ctx_set_level_loop": ([vp, i32], i32),
        "b200_ctx_forget_graph": ([vp], i32),
        "b200_ctx_set_sssp_delta": ([vp, C.c_float], i32),
        "b200_graph_no_in_arc_bitmap": ([vp, pg, vp], i32),
        "b200_graph_hot_columns": ([vp, pg, i64, vp, vp], i32),
        "b200_graph_first_in_neighbor": ([vp, pg, vp], i32),
        "b200_ctx_workspace": ([vp], vp),
        "b200_rmat_build_csr": ([vp, i32, i32, u64, vp, vp, vp, u64], i32),
        "b200_rmat_pairs": ([vp, i32, i32, u64, vp, vp], i32),
        "b200_build_csr_from_pairs": ([vp, i64, i64, vp, vp, i32, vp, vp, vp, u64], i32),
        "b200_advance_forward": ([vp, pg, pp, vp, i64, vp, i64, i32, i32, pi64, pi64], i32),
        "b200_filter": ([vp, pg, pp, vp, i64, vp, i64, i32, pi64], i32),
        "b200_uniquify": ([vp, pg, pp, vp, vp, i64, vp, i64, i32, pi64], i32),
        "b200_sparse_to_dense": ([vp, i64, vp, i64, vp], i32),
        "b200_dense_to_sparse": ([vp, i64, vp, vp, i64, pi64], i32),
        "b200_gen_unvisited": ([vp, pp, i64, vp, i64, pi64], i32),
        "b200_advance_backward": ([vp, pg, pp, vp, vp, i32, pi64, pi64], i32),
        "b200_neighborhood_reduce_f32": ([vp, pg, vp, i64, vp, vp, f32, i32, i32, i32, pi64], i32),
        "b200_bfs_run": ([vp, pg, i32, i32, f32, f32, vp, ps], i32),
        "b200_sssp_run": ([vp, pg, i32, vp, vp, ps], i32),
        "b200_pr_run": ([vp, pg, i32, i32, vp, vp, pi64, pi32, ps], i32),
        "b200_partition_locate": ([i32, i64, pi32, pi64], i32),
        "b200_partition_global_id": ([i32, i32, i64, pi64], i32),
        "b200_rmat_part_count": ([vp, i32, i32, u64, i32, i32, pi64], i32),
        "b200_rmat_build_csr_part": ([vp, i32, i32, u64, i32, i32, i64, vp, vp], i32),
        "b200_mg_bfs_init": ([vp, vp, i32, vp, pi64], i32),
        "b200_mg_bfs_push": ([vp, pg, vp, i32, vp, i64, vp, vp, i64, pi64, pi64, pi64], i32),
        "b200_mg_bfs_absorb": ([vp, pg, vp, i32, vp, i64, vp, pi64, pi64], i32),
        "b200_mg_bfs_pull": ([vp, pg, vp, i32, pi64, pi64, pi64], i32),
        "b200_mg_bitmap_or": ([vp, vp, vp, i64], i32),
        "b200_mg_list_to_slice": ([vp, vp, vp, i64, vp], i32),
        "b200_mg_slice_to_list": ([vp, vp, vp, vp, pi64], i32),
        "b200_p2p_bfs_heap_bytes": ([i32, i64, pi64], i32),
        "b200_p2p_bfs_create": ([vp, i32, i32, i64, C.POINTER(vp), vp, C.POINTER(vp)], i32),
        "b200_p2p_bfs_connect": ([vp, vp, vp], i32),
        "b200_p2p_bfs_run": ([vp, pg, i64, i32, i32, f32, f32, vp, ps, pi64], i32),
        "b200_p2p_bfs_prepare": ([vp, pg, i32, vp], i32),
        "b200_p2p_bfs_destroy": ([vp], i32),
        "b200_p2p_bfs_set_trace": ([vp, i32], i32),
        "b200_p2p_bfs_last_trace": ([vp, vp, i64, vp], i32),
        "b200_mtx_load": ([C.c_char_p, i32, C.POINTER(CHostCSR)], i32),
        "b200_csr_cache_write": ([C.c_char_p, C.POINTER(CHostCSR)], i32),
        "b200_csr_cache_read": ([C.c_char_p, C.POINTER(CHostCSR)], i32),
        "b200_host_csr_free": ([C.POINTER(CHostCSR)], i32),
        "b200_host_graph_upload": ([vp, i64, i64, vp, vp, vp, C.POINTER(vp)], i32),
        "b200_host_graph_free": ([vp, vp], i32),
        "b200_host_graph_view": ([vp, pg], i32),
        "b200_bfs_host": ([vp, vp, i32, i32, f32, f32, vp, vp, ps], i32),
        "b200_sssp_host": ([vp, vp, i32, vp, vp, vp, ps], i32),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)   # AttributeError => the library does not export what the header declares
        fn.argtypes = args
        fn.restype = res
    L._b200_signatures = sig
    _lib = L
    return L


def _check(status: int, what: str):
    if status != 0:
        raise B200Error(status, what, _lib.b200_last_cuda_error() if _lib is not None else 0)


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Stats:
    """Python view of b200_stats."""

    def __init__(self, c: CStats):
        self.num_levels = c.num_levels
        self.reached = c.reached
        self.total_arcs = c.total_arcs
        self.launches = c.launches
        self.device_ms = c.device_ms
        self.level_loop = "graph" if c.level_loop == LOOP_GRAPH else "host"
        self.levels = [
            dict(direction="pull" if l.direction else "push", frontier_len=l.frontier_len, arcs=l.arcs,
                 discovered=l.discovered, advance_ms=l.advance_ms, level_ms=l.level_ms,
                 exchange="bitmap" if l.reserved else "ids")   # peer-memory BFS: what crossed NVLink in this level
            for l in c.level[:min(c.num_levels, MAX_LEVELS)]
        ]


class Graph:
    """Device CSR (torch tensors own the memory).  Offsets are stored as int32 tensors holding
    uint32 bit patterns.  CSC aliases CSR (symmetric graphs; graph.hxx:75-80)."""

    def __init__(self, n: int, m: int, row_offsets, col_indices, col_values=None):
        self.n, self.m = int(n), int(m)
        self.row_offsets, self.col_indices, self.col_values = row_offsets, col_indices, col_values
        self.no_in_arc = None   # derived data (Context.prepare_graph), like graph_device_t::d_scanned_row_offsets
        self.first_in_nbr = None
        self.hot_ids = None     # derived data (Context.prepare_hot_columns)
        self.hot_indices = None
        self.csc = None         # (col_offsets, row_indices, row_values) of a directed graph's true transpose; None = aliases the CSR

    def set_csc(self, col_offsets, row_indices, row_values=None):
        """Attach the true CSC of a directed graph (d_col_offsets / d_row_indices / d_row_values, graph.hxx:44-46)."""
        self.csc = (col_offsets, row_indices, row_values)
        return self

    def cview(self) -> CGraph:
        co, ri, rv = self.csc if self.csc is not None else (self.row_offsets, self.col_indices, self.col_values)
        return CGraph(self.n, self.m, _ptr(self.row_offsets), _ptr(self.col_indices), _ptr(self.col_values),
                      _ptr(co), _ptr(ri), _ptr(rv), _ptr(self.no_in_arc),
                      _ptr(self.first_in_nbr), _ptr(self.hot_ids), _ptr(self.hot_indices),
                      0 if self.hot_ids is None else self.hot_ids.numel())

    def offsets_host(self):
        import numpy as np
        return self.row_offsets.cpu().numpy().view(np.uint32).astype(np.int64)

    def degrees_sum_reached(self, labels) -> int:
        """TEPS numerator: sum of deg(v) over reached v."""
        import torch
        off = self.row_offsets.to(torch.int64) & 0xFFFFFFFF
        deg = off[1:] - off[:-1]
        return int(deg[labels >= 0].sum().item())


def _quad_padded(n_elems: int, dtype, device):
    """Device array of n_elems whose allocation is readable up to the next multiple of 4 elements (b200_graph)."""
    import torch
    return torch.empty((max(n_elems, 1) + 3) // 4 * 4, dtype=dtype, device=device)[:n_elems]


class Context:
    """b200_ctx bound to one CUDA device; uses torch's current stream on that device by default."""

    def __init__(self, device: int = 0, stream: Optional[int] = "torch"):
        import torch
        self._L = load_library()
        self.device = device
        if not torch.cuda.is_available():
            raise RuntimeError("mini_b200 needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(device)
        if stream == "torch":
            # torch's default stream is the legacy default stream, whose handle is 0; pass the explicit
            # cudaStreamLegacy handle (0x1) so the ctx really shares it (NULL would make the ctx create a
            # private non-blocking stream that is NOT ordered with torch / NCCL work).
            stream = torch.cuda.current_stream(device).cuda_stream or 1
        h = C.c_void_p()
        _check(self._L.b200_ctx_create(C.byref(h), device, C.c_void_p(stream) if stream else None), "b200_ctx_create")
        self._h = h
        self.torch_device = torch.device("cuda", device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ misc
    def num_sms(self) -> int:
        v = C.c_int()
        _check(self._L.b200_ctx_num_sms(self._h, C.byref(v)), "b200_ctx_num_sms")
        return v.value

    def sync(self):
        _check(self._L.b200_ctx_sync(self._h), "b200_ctx_sync")

    def set_advance_impl(self, impl: int):
        """ADVANCE_QUAD (default; inside bfs / sssp the advance also creates the next level's scan), ADVANCE_QUAD_RESCAN
        (a scan kernel before every level) or ADVANCE_LBS: which kernel runs the push advance (same results)."""
        _check(self._L.b200_ctx_set_advance_impl(self._h, impl), "b200_ctx_set_advance_impl")

    def set_level_loop(self, impl: int):
        """LOOP_GRAPH (default: one CUDA graph per traversal, device-side decisions) or LOOP_HOST."""
        _check(self._L.b200_ctx_set_level_loop(self._h, impl), "b200_ctx_set_level_loop")

    def set_sssp_delta(self, delta: float):
        """First bucket width of the near-far order inside sssp(): 0 = automatic, float('inf') = the reference's
        Bellman-Ford frontier iterations.  Distances do not depend on it."""
        _check(self._L.b200_ctx_set_sssp_delta(self._h, float(delta)), "b200_ctx_set_sssp_delta")

    def l2_pin(self, tensor):
        if tensor is None:
            _check(self._L.b200_ctx_l2_pin(self._h, None, 0), "b200_ctx_l2_pin")
        else:
            _check(self._L.b200_ctx_l2_pin(self._h, tensor.data_ptr(), tensor.numel() * tensor.element_size()),
                   "b200_ctx_l2_pin")

    # ------------------------------------------------------------------ graphs
    def rmat_graph(self, scale: int, edge_factor: int = 16, seed: int = 1, weighted: bool = False,
                   weight_seed: int = 7) -> Graph:
        import torch
        n, m = 1 << scale, (2 * edge_factor) << scale
        off = torch.empty(n + 1, dtype=torch.int32, device=self.torch_device)
        idx = _quad_padded(m, torch.int32, self.torch_device)
        w = _quad_padded(m, torch.float32, self.torch_device) if weighted else None
        _check(self._L.b200_rmat_build_csr(self._h, scale, edge_factor, seed, off.data_ptr(), idx.data_ptr(),
                                           _ptr(w), weight_seed), "b200_rmat_build_csr")
        return Graph(n, m, off, idx, w)

    def rmat_pairs(self, scale: int, edge_factor: int = 16, seed: int = 1):
        import torch
        k = edge_factor << scale
        s = torch.empty(k, dtype=torch.int32, device=self.torch_device)
        d = torch.empty(k, dtype=torch.int32, device=self.torch_device)
        _check(self._L.b200_rmat_pairs(self._h, scale, edge_factor, seed, s.data_ptr(), d.data_ptr()), "b200_rmat_pairs")
        return s, d

    def csr_from_pairs(self, n: int, src, dst, symmetrize: bool = True, weighted: bool = False,
                       weight_seed: int = 7) -> Graph:
        import torch
        k = src.numel()
        m = k * (2 if symmetrize else 1)
        off = torch.empty(n + 1, dtype=torch.int32, device=self.torch_device)
        idx = _quad_padded(max(m, 1), torch.int32, self.torch_device)
        w = _quad_padded(max(m, 1), torch.float32, self.torch_device) if weighted else None
        _check(self._L.b200_build_csr_from_pairs(self._h, n, k, _ptr(src), _ptr(dst), int(symmetrize), off.data_ptr(),
                                                 idx.data_ptr(), _ptr(w), weight_seed), "b200_build_csr_from_pairs")
        return Graph(n, m, off, idx[:m], None if w is None else w[:m])

    def prepare_graph(self, g: Graph) -> Graph:
        """One-time derived data of a graph (b200_graph_no_in_arc_bitmap, b200_graph_first_in_neighbor): lets the
        pull levels of the direction-optimising BFS skip vertices without in-arcs and resolve most vertices from a
        contiguous array.  Optional -- without it the engine derives the bitmap per traversal and reads the CSC."""
        import torch
        bm = torch.empty((g.n + 31) // 32 + 1, dtype=torch.int32, device=self.torch_device)
        cg = g.cview()
        _check(self._L.b200_graph_no_in_arc_bitmap(self._h, C.byref(cg), bm.data_ptr()), "b200_graph_no_in_arc_bitmap")
        first = torch.empty(g.n, dtype=torch.int32, device=self.torch_device)
        _check(self._L.b200_graph_first_in_neighbor(self._h, C.byref(cg), first.data_ptr()), "b200_graph_first_in_neighbor")
        g.no_in_arc, g.first_in_nbr = bm, first
        return g

    def prepare_hot_columns(self, g: Graph, hot_count: int = 40960) -> Graph:
        """One-time derived data for neighbourhood reductions (b200_graph_hot_columns): the hot_count most frequent
        columns and the remapped index copy; the reduce kernel then serves their values from shared memory."""
        import torch
        hot_count = int(min(hot_count, g.n))
        ids = torch.empty(hot_count, dtype=torch.int32, device=self.torch_device)
        idx = _quad_padded(g.m, torch.int32, self.torch_device)
        cg = g.cview()
        _check(self._L.b200_graph_hot_columns(self._h, C.byref(cg), hot_count, ids.data_ptr(), idx.data_ptr()),
               "b200_graph_hot_columns")
        g.hot_ids, g.hot_indices = ids, idx
        return g

    def graph_from_host(self, offsets, indices, weights=None) -> Graph:
        """numpy CSR (any integer offsets < 2^32) -> device Graph."""
        import numpy as np
        import torch
        off = torch.from_numpy(np.asarray(offsets).astype(np.uint32).view(np.int32).copy()).to(self.torch_device)
        m = int(offsets[-1])
        idx = _quad_padded(m, torch.int32, self.torch_device)
        idx.copy_(torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int32)[:m]))
        w = None
        if weights is not None:
            w = _quad_padded(m, torch.float32, self.torch_device)
            w.copy_(torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float32)[:m]))
        return Graph(len(offsets) - 1, m, off, idx, w)

    def graph_from_mtx(self, path: str, undirected: bool = False) -> Graph:
        """MatrixMarket file -> device Graph, with the reference's load_graph semantics (b200_mtx_load)."""
        _n, off, idx, w = load_mtx(path, undirected)
        return self.graph_from_host(off, idx, w)

    # ------------------------------------------------------------------ primitives
    def bfs(self, g: Graph, src: int = 0, mode: int = BFS_PUSH, alpha: float = 0.0, beta: float = 0.0,
            labels=None, timing: bool = False):
        import torch
        if labels is None:
            labels = torch.empty(g.n, dtype=torch.int32, device=self.torch_device)
        cs = CStats()
        cs.collect_timing = int(timing)
        cg = g.cview()
        _check(self._L.b200_bfs_run(self._h, C.byref(cg), src, mode, alpha, beta, labels.data_ptr(), C.byref(cs)),
               "b200_bfs_run")
        return labels, Stats(cs)

    def sssp(self, g: Graph, src: int = 0, dist=None, preds=None, timing: bool = False):
        import torch
        if dist is None:
            dist = torch.empty(g.n, dtype=torch.float32, device=self.torch_device)
        cs = CStats()
        cs.collect_timing = int(timing)
        cg = g.cview()
        _check(self._L.b200_sssp_run(self._h, C.byref(cg), src, dist.data_ptr(), _ptr(preds), C.byref(cs)), "b200_sssp_run")
        return dist, Stats(cs)

    def pr(self, g: Graph, max_iter: int = 10, scatter: bool = False, timing: bool = False):
        import torch
        cur = torch.empty(g.n, dtype=torch.float32, device=self.torch_device)
        red = torch.empty(g.n, dtype=torch.float32, device=self.torch_device)
        lens = (C.c_int64 * max(max_iter, 1))()
        its = C.c_int()
        cs = CStats()
        cs.collect_timing = int(timing)
        cg = g.cview()
        _check(self._L.b200_pr_run(self._h, C.byref(cg), max_iter, int(scatter), cur.data_ptr(), red.data_ptr(), lens,
                                   C.byref(its), C.byref(cs)), "b200_pr_run")
        return cur, red, list(lens[:its.value]), Stats(cs)

    # ------------------------------------------------------------------ operators
    def advance_forward(self, g: Graph, problem: CProblem, frontier, out, iteration: int, flags: int = 0):
        n_out, arcs = C.c_int64(), C.c_int64()
        cg = g.cview()
        _check(self._L.b200_advance_forward(self._h, C.byref(cg), C.byref(problem), _ptr(frontier), frontier.numel(),
                                            _ptr(out), 0 if out is None else out.numel(), iteration, flags,
                                            C.byref(n_out), C.byref(arcs)), "b200_advance_forward")
        return n_out.value, arcs.value

    def filter(self, g: Optional[Graph], problem: CProblem, frontier, out, iteration: int) -> int:
        n_out = C.c_int64()
        cg = g.cview() if g is not None else None
        _check(self._L.b200_filter(self._h, C.byref(cg) if cg is not None else None, C.byref(problem), _ptr(frontier),
                                   frontier.numel(), _ptr(out), out.numel(), iteration, C.byref(n_out)), "b200_filter")
        return n_out.value

    def uniquify(self, g: Graph, problem: Optional[CProblem], visited_bitmap, frontier, out, iteration: int) -> int:
        n_out = C.c_int64()
        cg = g.cview()
        _check(self._L.b200_uniquify(self._h, C.byref(cg), C.byref(problem) if problem is not None else None,
                                     _ptr(visited_bitmap), _ptr(frontier), frontier.numel(), _ptr(out), out.numel(),
                                     iteration, C.byref(n_out)), "b200_uniquify")
        return n_out.value

    def sparse_to_dense(self, n: int, sparse, bitmap):
        _check(self._L.b200_sparse_to_dense(self._h, n, _ptr(sparse), sparse.numel(), _ptr(bitmap)), "b200_sparse_to_dense")

    def dense_to_sparse(self, n: int, bitmap, out) -> int:
        k = C.c_int64()
        _check(self._L.b200_dense_to_sparse(self._h, n, _ptr(bitmap), _ptr(out), out.numel(), C.byref(k)), "b200_dense_to_sparse")
        return k.value

    def gen_unvisited(self, problem: CProblem, n: int, out) -> int:
        k = C.c_int64()
        _check(self._L.b200_gen_unvisited(self._h, C.byref(problem), n, _ptr(out), out.numel(), C.byref(k)), "b200_gen_unvisited")
        return k.value

    def advance_backward(self, g: Graph, problem: CProblem, frontier_bitmap, next_bitmap, iteration: int):
        found, arcs = C.c_int64(), C.c_int64()
        cg = g.cview()
        _check(self._L.b200_advance_backward(self._h, C.byref(cg), C.byref(problem), _ptr(frontier_bitmap),
                                             _ptr(next_bitmap), iteration, C.byref(found), C.byref(arcs)),
               "b200_advance_backward")
        return found.value, arcs.value

    def neighborhood_reduce(self, g: Graph, frontier, values, reduced, identity: float = 0.0, op: int = OP_PLUS,
                            push: bool = False, scatter: bool = False) -> int:
        arcs = C.c_int64()
        cg = g.cview()
        _check(self._L.b200_neighborhood_reduce_f32(self._h, C.byref(cg), _ptr(frontier), frontier.numel(), _ptr(values),
                                                    _ptr(reduced), identity, op, int(push), int(scatter), C.byref(arcs)),
               "b200_neighborhood_reduce_f32")
        return arcs.value

    # ------------------------------------------------------------------ host-buffer entry points
    def host_graph_upload(self, offsets_u32, indices_i32, weights_f32=None):
        """numpy (ideally pinned-backed) CSR -> engine-owned device copy (graph_to_device, graph.hxx:60-83)."""
        h = C.c_void_p()
        n, m = offsets_u32.shape[0] - 1, indices_i32.shape[0]
        _check(self._L.b200_host_graph_upload(self._h, n, m, offsets_u32.ctypes.data, indices_i32.ctypes.data,
                                              None if weights_f32 is None else weights_f32.ctypes.data, C.byref(h)),
               "b200_host_graph_upload")
        return h

    def host_graph_free(self, h):
        self._L.b200_host_graph_free(self._h, h)

    def bfs_host(self, hg, src: int, labels_init_ptr: Optional[int], labels_out_ptr: int, mode: int = BFS_PUSH,
                 alpha: float = 0.0, beta: float = 0.0):
        cs = CStats()
        _check(self._L.b200_bfs_host(self._h, hg, src, mode, alpha, beta, labels_init_ptr, labels_out_ptr, C.byref(cs)),
               "b200_bfs_host")
        return Stats(cs)

    def sssp_host(self, hg, src: int, dist_init_ptr: Optional[int], dist_out_ptr: int, preds_out_ptr: Optional[int] = None):
        cs = CStats()
        _check(self._L.b200_sssp_host(self._h, hg, src, dist_init_ptr, dist_out_ptr, preds_out_ptr, C.byref(cs)),
               "b200_sssp_host")
        return Stats(cs)


def _copy_out(h: CHostCSR):
    import numpy as np
    n, m = h.n, h.m
    off = np.ctypeslib.as_array(C.cast(h.row_offsets, C.POINTER(C.c_uint32)), (n + 1,)).copy()
    idx = np.ctypeslib.as_array(C.cast(h.col_indices, C.POINTER(C.c_int32)), (max(m, 1),))[:m].copy()
    w = None
    if h.col_values:
        w = np.ctypeslib.as_array(C.cast(h.col_values, C.POINTER(C.c_float)), (max(m, 1),))[:m].copy()
    return n, off, idx, w


def load_mtx(path: str, undirected: bool = False):
    """b200_mtx_load (the reference's load_graph semantics) -> (n, offsets uint32 [n+1], indices int32 [m], weights f32 [m])."""
    L = load_library()
    h = CHostCSR()
    _check(L.b200_mtx_load(os.fsencode(path), int(undirected), C.byref(h)), "b200_mtx_load")
    try:
        return _copy_out(h)
    finally:
        L.b200_host_csr_free(C.byref(h))


def write_csr_cache(path: str, offsets_u32, indices_i32, weights_f32=None):
    import numpy as np
    L = load_library()
    off = np.ascontiguousarray(offsets_u32, dtype=np.uint32)
    idx = np.ascontiguousarray(indices_i32, dtype=np.int32)
    w = None if weights_f32 is None else np.ascontiguousarray(weights_f32, dtype=np.float32)
    h = CHostCSR(len(off) - 1, len(idx), off.ctypes.data, idx.ctypes.data, None if w is None else w.ctypes.data)
    _check(L.b200_csr_cache_write(os.fsencode(path), C.byref(h)), "b200_csr_cache_write")


def read_csr_cache(path: str):
    L = load_library()
    h = CHostCSR()
    _check(L.b200_csr_cache_read(os.fsencode(path), C.byref(h)), "b200_csr_cache_read")
    try:
        return _copy_out(h)
    finally:
        L.b200_host_csr_free(C.byref(h))


def bfs_problem(labels, visited_bitmap, preds=None) -> CProblem:
    p = CProblem()
    p.kind = PROBLEM_BFS
    p.labels, p.preds, p.visited_bitmap = _ptr(labels), _ptr(preds), _ptr(visited_bitmap)
    return p


def sssp_problem(dist, weights, visited, preds=None) -> CProblem:
    p = CProblem()
    p.kind = PROBLEM_SSSP
    p.dist, p.weights, p.visited, p.preds = _ptr(dist), _ptr(weights), _ptr(visited), _ptr(preds)
    return p


def pr_problem(current, reduced, degrees=None) -> CProblem:
    p = CProblem()
    p.kind = PROBLEM_PR
    p.current_ranks, p.reduced_ranks, p.degrees = _ptr(current), _ptr(reduced), _ptr(degrees)
    return p
