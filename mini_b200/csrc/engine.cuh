// engine.cuh -- private state behind the opaque b200_ctx handle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b200/workspace.h"
#include "b200_frontier.h"

namespace b200 { struct LevelLoop; struct NearFar; }

struct b200_ctx {
    b200_workspace ws;
    bool own_stream;
    // traversal scratch, sized by the largest n seen (ensure_traversal_scratch)
    int64_t scratch_n;
    int32_t *frontier[2];      // ping-pong sparse frontiers, n ints each (fused filter => |F| <= n)
    uint32_t *bm_visited;      // 1 bit / vertex
    uint32_t *bm_frontier[2];  // pull-phase frontier bitmaps
    int32_t *stamp;            // SSSP per-iteration dedupe stamps
    // per-level timing events (created lazily when stats->collect_timing)
    cudaEvent_t ev_run[2];
    cudaEvent_t *ev_level;     // 3 per level: level start, advance start, advance stop
    int ev_level_count;
    // L2 persistence
    bool l2_window_set;
    int adv_impl;              // B200_ADVANCE_QUAD | B200_ADVANCE_LBS
    int loop_impl;             // B200_LOOP_GRAPH | B200_LOOP_HOST
    b200::LevelLoop *loop;     // graph-driven level loop (level_loop.cu), created on first use
    b200::NearFar *d_nf;       // near-far SSSP state (near_far.cuh); d_nf_sum: scratch of the weight sample
    double *d_nf_sum;
    float sssp_delta;          // b200_ctx_set_sssp_delta: 0 = automatic, +inf = Bellman-Ford iterations, else the first bucket width
    const void *nf_key_w;      // the automatic width is computed once per (weights array, n, m)
    int64_t nf_key_n, nf_key_m;
    float nf_auto_delta;
    float *hot_vals;           // [B200_HOT_MAX] values of a graph's hot columns, refreshed by every hot neighbourhood reduce
    uint64_t scratch_gen;      // bumped whenever workspace / traversal scratch is reallocated: cached traversal graphs
                               // (level_loop.cu, p2p_bfs.cu) hold those addresses and are rebuilt when it changes
};

struct b200_host_graph {
    int64_t n, m;
    uint32_t *d_row_offsets;
    int32_t *d_col_indices;
    float *d_col_values;
    int32_t *d_labels;   // result staging (BFS)
    float *d_dist;       // result staging (SSSP)
    int32_t *d_preds;
    uint32_t *d_no_in_arc;   // b200_graph::no_in_arc_bitmap, built once at upload
    int32_t *d_first_in_nbr; // b200_graph::first_in_neighbor
};

namespace b200 {

extern thread_local int g_last_cuda_error;

inline int cuda_status(cudaError_t e) {
    if (e == cudaSuccess) return B200_OK;
    g_last_cuda_error = (int)e;
    (void)cudaGetLastError();   // clear the sticky-less error
    return e == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA;
}

#define B200_CUDA(call)                                   \
    do {                                                  \
        cudaError_t _e = (call);                          \
        if (_e != cudaSuccess) return ::b200::cuda_status(_e); \
    } while (0)

#define B200_TRY(call)                 \
    do {                               \
        int _s = (call);               \
        if (_s != B200_OK) return _s;  \
    } while (0)

int ensure_traversal_scratch(b200_ctx *ctx, int64_t n);

// The work-creating advance reads the row bounds of every vertex it emits inside its flush.  While the offsets
// array is L2-resident (scale 22: 16 MB) that is cheaper than a scan kernel between the levels (push BFS 237.6 ->
// 258 GTEPS); at scale 26 (268 MB of offsets, every read a DRAM round trip in the middle of a flush) it was 10 %
// slower than the streaming scan kernel (8.33 -> 9.13 ms), so large graphs keep the scan.
constexpr int64_t WORK_CREATE_MAX_N = 1ll << 23;
// Pull levels walk every unvisited vertex; on RMAT half of them have no arc at all (scale 26: 34 M of 67 M) and
// were re-inspected at every pull level.  At the push -> pull switch the visited bitmap of a direction-optimising
// traversal is therefore OR-ed with "vertex has no in-arc" (by the pull arrays): the pull kernel can never label
// such a vertex, so no result changes (labels stay -1), and whole words of them are skipped with one 4-byte read.
// Source: b200_graph::no_in_arc_bitmap if the caller built it once (b200_graph_no_in_arc_bitmap), else one pass over
// the pull offsets.  It is NOT applied before the switch: when the CSC arrays alias the CSR of a directed graph
// (graph.hxx:75-80 does that for every input) "no in-arc" really means "no out-arc", and the push levels must still
// label those sinks exactly as the reference does (bfs_enactor.hxx:54-66).
cudaError_t launch_no_in_arc_bitmap(b200_workspace *ws, const uint32_t *pull_offsets, int64_t n, uint32_t *d_bitmap);
cudaError_t launch_or_no_in_arc(b200_workspace *ws, const uint32_t *pull_offsets, int64_t n, const uint32_t *iso, uint32_t *d_visited);
// b200_graph::first_in_neighbor: out[v] = first in-neighbour of v, or -1 (see the pull kernel, advance.cuh)
cudaError_t launch_first_in_neighbor(b200_workspace *ws, const uint32_t *pull_offsets, const int32_t *pull_indices, int64_t n,
                                     int32_t *d_out);
cudaError_t preload_no_in_arc_kernel();   // lazy module loading vs. spinning peer kernels, see p2p_bfs.cu

// level_loop.cu
int bfs_run_graph(b200_ctx *ctx, const b200_graph *g, int32_t src, int mode, float alpha, float beta, int32_t *d_labels,
                  b200_stats *stats);
int sssp_run_graph(b200_ctx *ctx, const b200_graph *g, int32_t src, float *d_dist, float delta0, b200_stats *stats);
void level_loop_invalidate(b200_ctx *ctx);
void level_loop_destroy(b200_ctx *ctx);

}  // namespace b200
