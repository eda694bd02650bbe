/*
 * b200_frontier.h -- thin C ABI of the B200-native frontier-traversal engine.
 *
 * This is the drop-in boundary for the one hot path of gunrock/mini: the
 * advance / filter / neighborhood_reduce operators and the BFS / SSSP / PR-style
 * primitives built on them.  The reference has NO FFI: its boundary is the
 * header-only C++ template API in namespace gunrock::oprtr.  The
 * C++ mirror of that API lives in include/gunrock/ (same type and function
 * names, so tests/bfs/test_bfs.cu-style drivers compile unchanged) and calls
 * down into this C ABI; Python (ctypes), the tests and bench.py bind the same
 * symbols.  Each entry point cites the reference interface it replaces
 * (file:line relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; every d_* pointer is DEVICE memory owned
 *     by the caller, every h_* pointer is HOST memory owned by the caller;
 *   - every function returns an int status (B200_OK == 0); nothing throws or
 *     calls exit() across the ABI (the reference printf+exit(0)s on frontier
 *     overflow, frontier.hxx:53-59,84-89: here that is B200_ERR_OVERFLOW);
 *   - calls are synchronous at return with respect to their host outputs
 *     (the reference contract: enactors branch on the returned count,
 *     advance.hxx:43, kernel_compact.hxx:75-81) and are issued on the ctx
 *     stream; a ctx is not thread-safe (neither is standard_context_t);
 *   - row offsets are 4-byte UNSIGNED (bit-compatible with the reference's
 *     non-negative `int` offsets, graph.hxx:19-26, and able to hold RMAT
 *     scale-26's 2^31 arcs, which the reference cannot); vertex ids are int32.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point returns B200_ERR_CUDA.
 */
#ifndef B200_FRONTIER_H
#define B200_FRONTIER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 3

enum {
    B200_OK = 0,
    B200_ERR_CUDA = 1,      /* a CUDA runtime call failed: see b200_last_cuda_error() */
    B200_ERR_INVALID = 2,   /* bad argument */
    B200_ERR_OVERFLOW = 3,  /* an output frontier exceeded the capacity the caller gave */
    B200_ERR_NOMEM = 4,
    B200_ERR_UNSUPPORTED = 5,
    B200_ERR_TIMEOUT = 6,   /* a cross-GPU barrier of the peer-memory path waited too long for a peer */
    B200_ERR_IO = 7,        /* a file could not be opened, read or written */
    B200_ERR_FORMAT = 8     /* a file is not what it should be (malformed .mtx, bad / truncated / corrupt CSR cache) */
};

typedef struct b200_ctx b200_ctx; /* replaces mgpu::standard_context_t (context.hxx:103-219) + per-call mem_t scratch */

/* Device CSR / CSC view.  Mirrors gunrock::graph_device_t (graph.hxx:37-58);
 * col_offsets/row_indices/row_values may alias the CSR arrays exactly as
 * graph_to_device does for symmetric graphs (graph.hxx:75-80).
 * Index / value arrays that are 16-byte aligned are read as aligned 128-bit quads: their allocation must be
 * readable up to the next multiple of 4 elements (4 * ceil(m / 4)); the pad is never interpreted.  Every allocator
 * in this library (b200_host_graph_upload, the Python Graph builders) pads; arrays that are not 16-byte aligned take
 * the element-wise kernels.  Weights must be finite and >= 0 (distances are ordered through their bit patterns). */
typedef struct b200_graph {
    int64_t n;                   /* num_nodes */
    int64_t m;                   /* num_edges (arcs), < 2^32 */
    const uint32_t *row_offsets; /* d_row_offsets  [n+1] */
    const int32_t *col_indices;  /* d_col_indices  [m]   */
    const float *col_values;     /* d_col_values   [m] or NULL */
    const uint32_t *col_offsets; /* d_col_offsets  [n+1] (CSC) */
    const int32_t *row_indices;  /* d_row_indices  [m]   (CSC) */
    const float *row_values;     /* d_row_values   [m] or NULL */
    const uint32_t *no_in_arc_bitmap; /* optional derived data, like graph_device_t::d_scanned_row_offsets
                                    (graph.hxx:52): [(n+31)/32] words, bit v set iff vertex v has no in-arc, built
                                    once per graph by b200_graph_no_in_arc_bitmap; the direction-optimising BFS
                                    ORs it into its visited set at the push -> pull switch so pull levels skip vertices
                                    nobody can reach.  NULL => computed at that switch (one pass over the offsets) */
    const int32_t *first_in_neighbor; /* optional derived data: [n], the first in-neighbour of every vertex (row_indices[
                                    col_offsets[v]]) or -1, built once per graph by b200_graph_first_in_neighbor.  A pull
                                    level resolves most vertices at their first in-arc (scale-26 level 1: 28 M of 32 M);
                                    with this contiguous array they touch neither the offsets nor a random 32-byte
                                    sector of the index array.  NULL => the pull kernel reads the CSC itself */
    /* Optional derived data for neighbourhood reductions over the pull arrays (b200_graph_hot_columns, once per graph):
       hot_ids[hot_count] = the vertices that occur most often in row_indices, hot_indices[m] = a copy of row_indices in
       which every occurrence of hot_ids[k] is stored as ~k.  The reduce kernel keeps the hot vertices' values in shared
       memory, so their gathers (40 % of the arcs of an RMAT graph for 40 K hot vertices) never touch L1 / L2.  Results
       are identical; all NULL / 0 => plain row_indices. */
    const int32_t *hot_ids;
    const int32_t *hot_indices;
    int64_t hot_count;
} b200_graph;

/* Built-in problems = the reference's data_slice_t structs, as one POD.
 * kind selects the functor the operator entry points apply. */
enum { B200_PROBLEM_BFS = 1, B200_PROBLEM_SSSP = 2, B200_PROBLEM_PR = 3 };
typedef struct b200_problem {
    int32_t kind;
    int32_t reserved;
    /* BFS: bfs_problem_t::data_slice_t (bfs_problem.hxx:16-27) */
    int32_t *labels;        /* d_labels: -1 unvisited, else depth */
    int32_t *preds;         /* d_preds (BFS: never written by the reference; SSSP: racy last writer) */
    /* SSSP: sssp_problem_t::data_slice_t (sssp_problem.hxx:19-33) */
    float *dist;            /* d_labels (float): FLT_MAX unreached */
    const float *weights;   /* d_weights = gslice->d_col_values */
    int32_t *visited;       /* d_visited iteration stamps, init -1 */
    /* PR: pr_problem_t::data_slice_t (pr_problem.hxx:14-25) */
    float *current_ranks;
    float *reduced_ranks;
    float *degrees;
    /* engine-side visited bitmap (n/32 words, 1 bit per vertex), optional:
     * the "flexible uniquification" state of filter.hxx:33-119 done right */
    uint32_t *visited_bitmap;
} b200_problem;

/* Per-run statistics (new; the reference only prints wall clock, test_bfs.cu:38-42). */
#define B200_MAX_LEVELS 512
typedef struct b200_level_stat {
    int32_t direction;     /* 0 = push, 1 = pull */
    int32_t reserved;
    int64_t frontier_len;  /* |F_l| (push) or |U_l| (pull) */
    int64_t arcs;          /* m_l expanded (push) / in-arcs inspected (pull; counted in-kernel) */
    int64_t discovered;    /* |F_{l+1}| */
    float advance_ms;      /* device time of the advance (LBS or pull) kernel alone */
    float level_ms;        /* device time of the whole level (scan + advance + bookkeeping) */
} b200_level_stat;
typedef struct b200_stats {
    int32_t collect_timing; /* in: nonzero => record per-level CUDA-event timings (adds event overhead) */
    int32_t num_levels;     /* out */
    int64_t reached;        /* out: vertices reached */
    int64_t total_arcs;     /* out: sum of `arcs` over levels */
    int64_t launches;       /* out: kernels launched by this call */
    float device_ms;        /* out: device time source-in-frontier -> labels final */
    int32_t level_loop;     /* out: B200_LOOP_GRAPH or B200_LOOP_HOST -- which level loop produced this run */
    b200_level_stat level[B200_MAX_LEVELS];
} b200_stats;

/* ---- library / context ---------------------------------------------------- */
int b200_abi_version(void);
const char *b200_status_string(int status);
int b200_last_cuda_error(void);                 /* cudaError_t of the last failing call on this thread */
int b200_device_count(int *count);

/* device: CUDA ordinal.  stream: a cudaStream_t to issue on, or NULL to let the
 * ctx create its own non-blocking stream.  Replaces `standard_context_t context;`
 * (test_bfs.cu:22; context.hxx:103-110). */
int b200_ctx_create(b200_ctx **out, int device, void *stream);
int b200_ctx_destroy(b200_ctx *ctx);
/* Pre-size the scratch arena so no operator call allocates (the reference
 * cudaMalloc/cudaFree's 3-6 times per operator call). */
int b200_ctx_reserve(b200_ctx *ctx, int64_t max_frontier_items);
int b200_ctx_sync(b200_ctx *ctx);
int b200_ctx_num_sms(b200_ctx *ctx, int *num_sms);
/* Keep [ptr, ptr+bytes) L2-resident for kernels on the ctx stream
 * (cudaAccessPolicyWindow, persisting hits).  bytes = 0 clears the window. */
int b200_ctx_l2_pin(b200_ctx *ctx, const void *d_ptr, int64_t bytes);
/* Which kernel implements the push advance (transform_scan + transform_lbs, advance.hxx:20-67)
 * behind the operators and primitives below.  Both give identical results.
 *   B200_ADVANCE_QUAD (default): LBS over aligned 16-byte quads of col_indices, warp-private
 *     chunks, probe / compact / dense-commit stages, L1-resident bitmap (include/b200/quad_advance.cuh);
 *     Inside b200_bfs_run / b200_sssp_run it is WORK-CREATING (the README's "dynamic group workload mapping",
 *     README.md:12; nearest in-tree relative: mgpu expt::lbs_workcreate, kernel_workcreate.hxx:15-268): the flush
 *     that appends a discovered vertex to the next frontier also appends its row bounds and its position in the
 *     next level's quad scan -- slots and scan positions come from ONE packed 64-bit atomic, so the scan stays
 *     sorted -- and the next level starts without a scan kernel (scale-22 push BFS 0.571 -> 0.526 ms; the heaviest
 *     launch takes 0.2411 instead of 0.2297 ms for 5.5 % more bytes, profiles/README.md).  Graphs with more than
 *     2^23 vertices keep the scan: their offsets no longer fit in L2 and the flush's reads become DRAM round trips
 *     (scale 26: 10 % slower than the streaming scan kernel);
 *   B200_ADVANCE_QUAD_RESCAN: the same kernel with a degree scan before every level (transform_scan + transform_lbs
 *     as the reference orders them, advance.hxx:30-44);
 *   B200_ADVANCE_LBS: LBS over arcs, CTA-cooperative windows (include/b200/advance.cuh); also
 *     taken automatically for B200_ADV_RAW_OUTPUT and for arrays that are not 16-byte aligned. */
enum { B200_ADVANCE_QUAD = 0, B200_ADVANCE_LBS = 1, B200_ADVANCE_QUAD_RESCAN = 2 };
int b200_ctx_set_advance_impl(b200_ctx *ctx, int impl);

/* Who drives the level loop of b200_bfs_run (the role of the host loop in bfs_enactor_t::enact_pushpull,
 * bfs_enactor.hxx:54-115, which reads a count back from the device after every operator):
 *   B200_LOOP_GRAPH (default): the whole traversal is one CUDA graph; WHILE / IF / SWITCH conditional nodes are
 *     set by a device-side decide kernel that applies the same push / pull / stop rules, so the host
 *     synchronises once per traversal (mini_b200/csrc/level_loop.cu).  Taken when the quad advance is
 *     usable and no per-level timing is requested; b200_stats.level_loop reports what ran;
 *   B200_LOOP_HOST: one counter read-back and host decision per level (also the timing path). */
enum { B200_LOOP_GRAPH = 0, B200_LOOP_HOST = 1 };
int b200_ctx_set_level_loop(b200_ctx *ctx, int impl);

/* The graph-driven level loop caches the instantiated traversal graph keyed by the addresses of the b200_graph
 * arrays and of the labels buffer; array CONTENTS are read afresh by every traversal, so rewriting a graph in place
 * needs nothing.  This drops the cached traversal graph (e.g. before freeing the buffers it references). */
int b200_ctx_forget_graph(b200_ctx *ctx);

/* Fills d_bitmap[(n+31)/32] for b200_graph::no_in_arc_bitmap (bit v set iff in-degree(v) == 0, from col_offsets --
 * the CSC the pull advance walks, advance.hxx:108-160).  One-time per graph, beside graph_to_device (graph.hxx:60-83). */
int b200_graph_no_in_arc_bitmap(b200_ctx *ctx, const b200_graph *g, uint32_t *d_bitmap);
/* Fills b200_graph::hot_ids / hot_indices for g's pull arrays: d_hot_ids[hot_count] receives the hot_count vertices with
 * the most occurrences in row_indices (ties: smaller id first), d_hot_indices[m] the remapped index copy (readable up to
 * the next multiple of 4 elements, like every index array).  hot_count <= B200_HOT_MAX. */
#define B200_HOT_MAX 40960
int b200_graph_hot_columns(b200_ctx *ctx, const b200_graph *g, int64_t hot_count, int32_t *d_hot_ids, int32_t *d_hot_indices);
/* Fills d_out[n] for b200_graph::first_in_neighbor.  One-time per graph. */
int b200_graph_first_in_neighbor(b200_ctx *ctx, const b200_graph *g, int32_t *d_out);

/* ---- synthetic input (the reference has only load_graph, graph.hxx:96-223) */
/* Symmetrised RMAT(0.57,0.19,0.19,0.05): n = 2^scale, m = 2*edge_factor*2^scale arcs,
 * duplicates and self loops kept, arcs sorted by (src,dst) -- what graph.hxx:139-172 builds.
 * d_weights (nullable): integer weights in [1,64] per undirected pair, as exact floats. */
int b200_rmat_build_csr(b200_ctx *ctx, int scale, int edge_factor, uint64_t seed,
                        uint32_t *d_row_offsets /* [n+1] */, int32_t *d_col_indices /* [m] */,
                        float *d_weights /* [m] or NULL */, uint64_t weight_seed);
/* Same pairs as the above, un-symmetrised, for generator parity tests: d_src/d_dst [edge_factor<<scale] */
int b200_rmat_pairs(b200_ctx *ctx, int scale, int edge_factor, uint64_t seed, int32_t *d_src, int32_t *d_dst);
/* CSR from an arbitrary device pair list (symmetrize adds every reverse); same ordering rules. */
int b200_build_csr_from_pairs(b200_ctx *ctx, int64_t n, int64_t npairs, const int32_t *d_src,
                              const int32_t *d_dst, int symmetrize, uint32_t *d_row_offsets,
                              int32_t *d_col_indices, float *d_weights, uint64_t weight_seed);

/* ---- operators with the built-in functors --------------------------------- */
enum { B200_ADV_IDEMPOTENT = 1, B200_ADV_NO_OUTPUT = 2, B200_ADV_RAW_OUTPUT = 4 };

/* advance_forward_kernel<Problem,Functor,idempotence,has_output> (advance.hxx:20-67).
 * Expands the out-arcs of d_in[0..in_len) with a merge-path load-balanced search and
 * applies the problem's cond_advance/apply_advance (bfs_functor.hxx:26-33,
 * sssp_functor.hxx:20-34).  Default output is the COMPACTED list of accepted
 * neighbours (advance + the `!= -1` filter fused, written with CTA-aggregated
 * appends); B200_ADV_RAW_OUTPUT reproduces the reference layout out[idx] = nbr | -1
 * for idx in [0, m_F).  *out_len = items written, *arcs = m_F. */
int b200_advance_forward(b200_ctx *ctx, const b200_graph *g, const b200_problem *p,
                         const int32_t *d_in, int64_t in_len, int32_t *d_out, int64_t out_capacity,
                         int iteration, int flags, int64_t *out_len, int64_t *arcs);

/* filter_kernel<Problem,Functor> (filter.hxx:11-31): stable compaction of the items
 * for which cond_filter holds (bfs_functor.hxx:9-11, sssp_functor.hxx:12-18,
 * pr_functor.hxx:11-17; predicates run exactly once per item, side effects included). */
int b200_filter(b200_ctx *ctx, const b200_graph *g, const b200_problem *p, const int32_t *d_in,
                int64_t in_len, int32_t *d_out, int64_t out_capacity, int iteration, int64_t *out_len);

/* uniquify_kernel<Problem,ProblemFunctor> (filter.hxx:95-119): drops -1 and every
 * vertex already in the visited bitmap (1 bit per vertex, [ceil(n/32)] words, exact
 * test-and-set instead of the reference's heuristic culls), marks survivors and,
 * for BFS, labels them iteration+1 (bfs_functor.hxx:13-24).  Correct for any source. */
int b200_uniquify(b200_ctx *ctx, const b200_graph *g, const b200_problem *p, uint32_t *d_visited_bitmap,
                  const int32_t *d_in, int64_t in_len, int32_t *d_out, int64_t out_capacity,
                  int iteration, int64_t *out_len);

/* sparse_to_dense_kernel (advance.hxx:69-84): frontier list -> 1-bit-per-vertex bitmap
 * (the reference uses one int per vertex, bfs_enactor.hxx:83). bitmap is cleared first. */
int b200_sparse_to_dense(b200_ctx *ctx, int64_t n, const int32_t *d_sparse, int64_t len, uint32_t *d_bitmap);
/* inverse; order is ascending vertex id. */
int b200_dense_to_sparse(b200_ctx *ctx, int64_t n, const uint32_t *d_bitmap, int32_t *d_sparse,
                         int64_t capacity, int64_t *out_len);
/* gen_unvisited_kernel (advance.hxx:86-106) for BFS: {v : labels[v] == -1}, ascending. */
int b200_gen_unvisited(b200_ctx *ctx, const b200_problem *p, int64_t n, int32_t *d_unvisited,
                       int64_t capacity, int64_t *out_len);
/* advance_backward_kernel (advance.hxx:108-160), BFS pull step over a bitmap frontier:
 * every vertex with visited bit clear scans its in-arcs (CSC) and stops at the first
 * in-neighbour whose bit is set in d_frontier_bitmap; it is then labelled iteration+1,
 * set in d_next_bitmap and in the visited bitmap.  *discovered = |F_next|,
 * *arcs_inspected = in-arcs actually read. */
int b200_advance_backward(b200_ctx *ctx, const b200_graph *g, const b200_problem *p,
                          const uint32_t *d_frontier_bitmap, uint32_t *d_next_bitmap, int iteration,
                          int64_t *discovered, int64_t *arcs_inspected);

/* neighborhood_kernel<Problem,Functor,Value,reduce_op,has_output,push> (neighborhood.hxx:12-70)
 * for Value = float: d_reduced[slot] = op over the (push: out / pull: in) neighbours u of
 * d_in[slot] of d_values[u] (non-finite values read as 0, pr_functor.hxx:27-29); identity
 * when the neighbourhood is empty and never folded into non-empty ones
 * (mgpu tests/test_segreduce.cu:40-58).  scatter != 0 writes d_reduced[vertex] instead
 * (write_reduced_value, neighborhood.hxx:60-67).  fp32 sum order is unspecified:
 * callers compare within 1e-5 * sum|terms| + 1e-6. */
enum { B200_OP_PLUS = 0, B200_OP_MIN = 1, B200_OP_MAX = 2 };
int b200_neighborhood_reduce_f32(b200_ctx *ctx, const b200_graph *g, const int32_t *d_in, int64_t in_len,
                                 const float *d_values, float *d_reduced, float identity, int op,
                                 int push, int scatter, int64_t *arcs);

/* ---- whole primitives (device-resident inputs) ----------------------------- */
enum { B200_BFS_PUSH = 0,       /* reference default: alpha = 1/n never switches (test_bfs.cu:30) */
       B200_BFS_REF_ALPHA = 1,  /* bfs_enactor.hxx:68: one-way push->pull when unvisited < |F|*alpha */
       B200_BFS_BEAMER = 2 };   /* direction-optimising both ways: m_F > m_unvisited/alpha -> pull; |F| < n/beta -> push */

/* bfs_enactor_t::enact_pushpull (bfs_enactor.hxx:41-117) + bfs_problem_t ctor
 * (bfs_problem.hxx:34-46: labels = -1, labels[src] = 0).  d_labels [n] is fully overwritten. */
int b200_bfs_run(b200_ctx *ctx, const b200_graph *g, int32_t src, int mode, float alpha, float beta,
                 int32_t *d_labels, b200_stats *stats /* nullable */);

/* sssp_enactor_t::enact (sssp_enactor.hxx:40-72) + sssp_problem_t ctor
 * (sssp_problem.hxx:40-52: labels = FLT_MAX, labels[src] = 0, preds = -1).
 * d_dist [n] overwritten; d_preds nullable. */
int b200_sssp_run(b200_ctx *ctx, const b200_graph *g, int32_t src, float *d_dist, int32_t *d_preds,
                  b200_stats *stats /* nullable */);
/* Order of the frontier iterations inside b200_sssp_run.  The reference expands every improved vertex
 * in the next iteration (Bellman-Ford frontiers, sssp_enactor.hxx:51-70); b200_sssp_run holds back vertices whose new
 * distance is >= a moving cutoff until every smaller distance is final ("near-far": buckets of width delta taken from
 * the distance array itself, mini_b200/csrc/near_far.cuh), which relaxes ~1.0x instead of ~1.8x the reached arcs on
 * RMAT-22.  Distances are identical for any width (the same fixed point); weights must be non-negative.
 *   delta = 0 (default): 3.5 * mean weight / mean degree, sampled once per graph;  delta = INFINITY: the reference's
 *   order;  else: the first bucket width.  stats->num_levels counts advance iterations in every case. */
int b200_ctx_set_sssp_delta(b200_ctx *ctx, float delta);

/* pr_enactor_t::enact (pr_enactor.hxx:41-79) + pr_problem_t ctor (pr_problem.hxx:34-44).
 * scatter = 0 reproduces the reference's slot-indexed reduced[],
 * scatter = 1 is the intended vertex-indexed form.  d_current/d_reduced [n] overwritten;
 * h_frontier_lens (nullable, max_iter entries) gets each iteration's output length. */
int b200_pr_run(b200_ctx *ctx, const b200_graph *g, int max_iter, int scatter, float *d_current,
                float *d_reduced, int64_t *h_frontier_lens, int *iterations, b200_stats *stats);

/* ---- multi-GPU BFS: per-rank steps (new; the reference is single-GPU by design, README.md:4) -------
 * Swizzled-cyclic 1D vertex partition over P = 2^k <= 8 ranks (include/b200/partition.cuh): vertex v is
 * row v >> k of its owner's CSR (column ids stay global) and owner(v) = (v & (P-1)) ^ swizzle(v >> k),
 * swizzle(r) = top k bits of (r * 0x9E3779B1 mod 2^32) -- plain cyclic ownership gives rank 0 of 8
 * 44 % of an un-permuted RMAT graph's arcs, the swizzle is within 0.2 % of m/P.  Bitmaps are indexed
 * rank-major: bit(v) = owner(v) * n_local + (v >> k), so a rank's slice is contiguous and an all-gather
 * of the slices rebuilds the whole bitmap.  One process drives one GPU; between the calls below the host
 * layer issues the NCCL collectives (alltoall of counts + alltoallv of vertex ids after a push
 * level, allgather of frontier-bitmap slices before a pull level). */
/* The partition map itself (host arithmetic, no GPU): owner / local row of v, and the inverse. */
int b200_partition_locate(int num_ranks, int64_t v, int32_t *owner, int64_t *row);
int b200_partition_global_id(int num_ranks, int32_t rank, int64_t row, int64_t *v);
typedef struct b200_mg_bfs_state {
    int32_t rank, num_ranks;
    int64_t n_global, n_local;        /* n_local = n_global / P, a multiple of 32 */
    int32_t *labels;                  /* [n_local] local depths */
    uint32_t *known;                  /* [n_global/32] visited (owned vertices) / already-sent (others) */
    uint32_t *frontier_bitmap;        /* [n_global/32] all-gathered frontier of the current pull level */
    uint32_t *next_slice;             /* [n_local/32] this rank's slice of the next frontier bitmap */
    unsigned long long *box_counts;   /* [8] device counters of the per-destination boxes */
} b200_mg_bfs_state;

int b200_rmat_part_count(b200_ctx *ctx, int scale, int edge_factor, uint64_t seed, int rank, int num_ranks,
                         int64_t *m_local);
/* Local CSR of `rank`: n_local rows, global column ids, same ordering rules as b200_rmat_build_csr. */
int b200_rmat_build_csr_part(b200_ctx *ctx, int scale, int edge_factor, uint64_t seed, int rank, int num_ranks,
                             int64_t m_local, uint32_t *d_row_offsets /* [n_local+1] */,
                             int32_t *d_col_indices /* [m_local] */);
int b200_mg_bfs_init(b200_ctx *ctx, const b200_mg_bfs_state *s, int32_t src, int32_t *d_frontier,
                     int64_t *frontier_len);
/* One push level over the local frontier (global ids, all owned).  Local discoveries are labelled and
 * appended to d_next_frontier, remote ones to d_send_boxes[owner] (bucketed inside the advance
 * kernel).  h_counts[p] = items in box p (h_counts[rank] = local discoveries). */
int b200_mg_bfs_push(b200_ctx *ctx, const b200_graph *g_local, const b200_mg_bfs_state *s, int level,
                     const int32_t *d_frontier, int64_t frontier_len, int32_t *d_next_frontier,
                     int32_t *const *d_send_boxes /* host array [P] of device pointers */, int64_t box_capacity,
                     int64_t *h_counts /* [P] */, int64_t *arcs, int64_t *next_degree);
/* Receiver side: label-if-unvisited over the vertices peers sent; survivors join d_next_frontier.
 * *next_len = total length of the next frontier so far. */
int b200_mg_bfs_absorb(b200_ctx *ctx, const b200_graph *g_local, const b200_mg_bfs_state *s, int level,
                       const int32_t *d_inbox, int64_t count, int32_t *d_next_frontier, int64_t *next_len,
                       int64_t *next_degree);
/* One pull level over the local rows against s->frontier_bitmap; writes s->next_slice. */
int b200_mg_bfs_pull(b200_ctx *ctx, const b200_graph *g_local, const b200_mg_bfs_state *s, int level,
                     int64_t *found, int64_t *arcs_inspected, int64_t *found_degree);
int b200_mg_bitmap_or(b200_ctx *ctx, uint32_t *d_dst, const uint32_t *d_src, int64_t words);
int b200_mg_list_to_slice(b200_ctx *ctx, const b200_mg_bfs_state *s, const int32_t *d_list, int64_t len,
                          uint32_t *d_slice);
int b200_mg_slice_to_list(b200_ctx *ctx, const b200_mg_bfs_state *s, const uint32_t *d_slice, int32_t *d_list,
                          int64_t *len);

/* ---- multi-GPU BFS over NVLink PEER MEMORY: the exchange fused into the kernels (new) ----
 * Same partition and algorithm as the b200_mg_* steps, but no collective library on the data path:
 * every rank owns a symmetric heap (control block, two frontier-bitmap slices, one inbox segment per
 * sender) that its peers map (CUDA IPC between processes, plain pointers inside one process);
 *   push level : the advance kernel's flush stores remote discoveries straight into the owner's inbox
 *                (the alltoallv), a flag barrier publishes the counts, one absorb kernel labels them;
 *   pull level : one kernel reads all peers' bitmap slices through NVLink and folds them into `known`
 *                (the allgather), then the early-exit pull runs over the local rows;
 *   every level: the {|F_next|, arcs, deg, sent} rows are written into every peer's heap, summed
 *                after the barrier (the allreduce) and read by the host from mapped pinned memory.
 * All ranks must call b200_p2p_bfs_run with the same (src, mode, alpha, beta); it returns after the
 * whole traversal.  B200_ERR_TIMEOUT if a peer never arrives at a barrier (20 s). */
#define B200_IPC_HANDLE_BYTES 64
typedef struct b200_p2p_bfs b200_p2p_bfs;
int b200_p2p_bfs_heap_bytes(int num_ranks, int64_t n_global, int64_t *bytes);
/* n_global / num_ranks must be a multiple of 128.  ipc_handle_out (nullable) receives the
 * cudaIpcMemHandle_t of the heap (B200_IPC_HANDLE_BYTES); heap_base_out (nullable) its address. */
int b200_p2p_bfs_create(b200_ctx *ctx, int rank, int num_ranks, int64_t n_global, b200_p2p_bfs **out,
                        void *ipc_handle_out, void **heap_base_out);
/* Map the peers' heaps: ipc_handles = num_ranks consecutive handles gathered from all ranks (other
 * processes), or peer_bases = heap addresses valid in this process (ranks as threads / streams). */
int b200_p2p_bfs_connect(b200_p2p_bfs *s, const void *ipc_handles, void *const *peer_bases);
int b200_p2p_bfs_run(b200_p2p_bfs *s, const b200_graph *g_local, int64_t m_global, int32_t src, int mode,
                     float alpha, float beta, int32_t *d_labels_local, b200_stats *stats /* nullable */,
                     int64_t *sent_per_level /* nullable, [B200_MAX_LEVELS] vertices all ranks sent */);
/* Optional: build (and upload) the traversal graph of the graph-driven level loop for (g, mode, labels buffer)
 * ahead of the first b200_p2p_bfs_run, so that no rank is still instantiating a graph while its peers already
 * spin in a flag barrier (ranks that are threads of ONE process share a CUDA context).  A no-op when the ctx
 * uses B200_LOOP_HOST.  b200_p2p_bfs_run builds lazily if this was not called. */
int b200_p2p_bfs_prepare(b200_p2p_bfs *s, const b200_graph *g_local, int mode, int32_t *d_labels_local);
/* Timeline of the graph-driven loop (new; with no host between the levels there are no CUDA events to hang a
 * per-level time on): when on, the kernels of a run log (globaltimer ns << 8 | id) at their phase boundaries; ids
 * 64 + (level & 63) mark "level closed".  b200_p2p_bfs_last_trace copies the entries of the last traced run.
 * Other ids (profiles/format_trace.py names them): 20-29 small-level kernel phases, 2 / 3 / 5 / 13 / 9 / 12 big push level
 * (scan, claim-only advance, bitmap absorb, its flag barrier, stats + decide, its flag barrier), 30-36 pull-levels kernel
 * phases.  The single-GPU loops (b200_bfs_run / b200_sssp_run) print the same kind of line to stderr when the environment
 * variable B200_LOOP_TRACE is set (ids 2 scan, 3 advance, 9 decide, 11 bitmap -> list, 40 / 41 near-far passes). */
int b200_p2p_bfs_set_trace(b200_p2p_bfs *s, int on);
int b200_p2p_bfs_last_trace(b200_p2p_bfs *s, uint64_t *entries, int64_t capacity, int64_t *count);
int b200_p2p_bfs_destroy(b200_p2p_bfs *s);

/* ---- graph ingest on the host (no GPU involved): the .mtx loader of the reference and a binary CSR cache ---- */
/* Host CSR owned by the library (malloc; release with b200_host_csr_free). */
typedef struct b200_host_csr {
    int64_t n, m;
    uint32_t *row_offsets; /* [n+1] */
    int32_t *col_indices;  /* [m]   */
    float *col_values;     /* [m] or NULL */
} b200_host_csr;
/* load_graph (graph.hxx:96-223) for hosts that cannot include the C++ headers: "%" comment lines, a "rows cols
 * entries" line, then 1-based "i j [w]" entries; entry (i, j) is the arc (j-1) -> (i-1) with weight w (1.0 if absent);
 * `undirected` appends every reverse; arcs ordered by (row, column) with a STABLE sort (the reference's comparator is
 * not a strict weak order, :139-157), duplicates and self loops kept.  The result feeds b200_host_graph_upload.
 * B200_ERR_IO: cannot open; B200_ERR_FORMAT: malformed (the reference prints and exit(0)s, :108-111,:121-124). */
int b200_mtx_load(const char *path, int undirected, b200_host_csr *out);
/* Binary cache of a host CSR: 32-byte header (magic "B200CSR1", n, m, flags), the three arrays, FNV-1a checksum.
 * Reading verifies magic, sizes, checksum, offset monotonicity and index range (B200_ERR_FORMAT otherwise). */
int b200_csr_cache_write(const char *path, const b200_host_csr *csr);
int b200_csr_cache_read(const char *path, b200_host_csr *out);
int b200_host_csr_free(b200_host_csr *csr);

/* ---- host-buffer entry points (what test_bfs.cu times + extract: H2D, run, D2H) */
typedef struct b200_host_graph b200_host_graph; /* graph_to_device result (graph.hxx:60-83) kept by the engine */
int b200_host_graph_upload(b200_ctx *ctx, int64_t n, int64_t m, const uint32_t *h_row_offsets,
                           const int32_t *h_col_indices, const float *h_col_values /* nullable */,
                           b200_host_graph **out);
int b200_host_graph_free(b200_ctx *ctx, b200_host_graph *hg);
int b200_host_graph_view(const b200_host_graph *hg, b200_graph *out);
/* bfs_problem_t(graph, src) [H2D of the initial labels] + enact_pushpull + extract() [D2H]:
 * h_labels_init (nullable => engine builds labels on device) / h_labels_out are pinned or pageable host arrays. */
int b200_bfs_host(b200_ctx *ctx, b200_host_graph *hg, int32_t src, int mode, float alpha, float beta,
                  const int32_t *h_labels_init, int32_t *h_labels_out, b200_stats *stats);
int b200_sssp_host(b200_ctx *ctx, b200_host_graph *hg, int32_t src, const float *h_dist_init,
                   float *h_dist_out, int32_t *h_preds_out, b200_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* B200_FRONTIER_H */
