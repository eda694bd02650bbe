// engine.cu -- C-ABI operators and primitives with the built-in functors.
// Each entry point cites the reference interface it replaces in
// include/b200_frontier.h.  All kernels come from include/b200/*.cuh.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include "b200/operators.cuh"
#include "engine.cuh"
#include "near_far.cuh"

using namespace b200;

namespace {

// ----------------------------------------------------------------- small kernels
__global__ void bfs_init_kernel(int32_t *labels, uint32_t *visited, int32_t *frontier, int src) {
    labels[src] = 0;
    visited[src >> 5] |= 1u << (src & 31);
    frontier[0] = src;
}
__global__ void set_counter_kernel(unsigned long long *p, unsigned long long v) { *p = v; }
__global__ void sssp_init_kernel(float *dist, int32_t *preds, int32_t *stamp, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        dist[i] = FLT_MAX;
        if (preds) preds[i] = -1;
        stamp[i] = -1;
    }
}
__global__ void sssp_seed_kernel(float *dist, int32_t *frontier, int src) {
    dist[src] = 0.0f;
    frontier[0] = src;
}
__global__ void iota_fill_kernel(int32_t *frontier, int32_t *preds, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        frontier[i] = (int32_t)i;
        preds[i] = INT_MAX;
    }
}
__global__ void preds_fixup_kernel(int32_t *preds, unsigned long long n, int src) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (preds[i] == INT_MAX || i == (unsigned long long)src) preds[i] = -1;
}
__global__ void pr_init_kernel(float *current, float *reduced, int32_t *frontier, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        current[i] = 0.15f;
        reduced[i] = 0.0f;
        frontier[i] = (int32_t)i;
    }
}
// ----------------------------------------------------------------- compaction predicates
// cond_filter of bfs_functor_t (bfs_functor.hxx:9-11)
struct BfsFilterPred {
    const int *in;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = in[idx];
        return item != -1;
    }
};
// cond_filter of sssp_functor_t (sssp_functor.hxx:12-18); the stamp test-and-set is made atomic
struct SsspFilterPred {
    const int *in;
    int *stamp;
    int iteration;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = in[idx];
        if (item == -1) return false;
        return atomicExch(stamp + item, iteration) != iteration;
    }
};
// cond_filter of pr_functor_t (pr_functor.hxx:11-17); degrees == NULL => computed from the offsets
struct PrFilterPred {
    const int *in;
    float *current;
    const float *reduced;
    const float *degrees;
    const uint32_t *offsets;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = in[idx];
        const float old_value = current[item];
        const float deg = degrees ? degrees[item] : (float)(offsets[item + 1] - offsets[item]);
        float new_value = (deg > 0) ? (0.15f + 0.85f * reduced[item] / deg) : 0.15f;
        if (!isfinite(new_value)) new_value = 0;
        current[item] = new_value;
        return fabsf(new_value - old_value) > (0.001f * old_value);
    }
};
// uniquify (filter.hxx:95-119) with an exact bitmap test-and-set; BFS also labels (bfs_functor.hxx:13-24)
struct UniquifyPred {
    const int *in;
    uint32_t *visited;
    int *labels;   // nullable
    int label;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = in[idx];
        if (item < 0) return false;
        const uint32_t bit = 1u << (item & 31);
        if ((visited[item >> 5] & bit) || (atomicOr(visited + (item >> 5), bit) & bit)) return false;
        if (labels) labels[item] = label;
        return true;
    }
};
struct UnvisitedPred {   // cond_gen_unvisited (bfs_functor.hxx:39-41) over iota(n)
    const int *labels;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = (int)idx;
        return labels[idx] == -1;
    }
};
struct BitmapPred {
    const uint32_t *bitmap;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = (int)idx;
        return (bitmap[idx >> 5] >> (idx & 31)) & 1u;
    }
};

// Deterministic predecessors once the distances are final: the smallest u != v with
// dist[u] + w(u,v) == dist[v]; zero-weight arcs count only from the smaller id to the larger (no cycles,
// see SsspPredQ).  (The reference writes preds[dst] = src from every relaxation attempt,
// sssp_functor.hxx:31-34, so its GPU preds are a race)
struct SsspPredOp {
    const float *dist;
    const float *weights;
    int *preds;
    __device__ __forceinline__ bool probe(int src, int dst, uint32_t eid) const {
        const float ds = dist[src], w = ld_stream(weights + eid);
        return ds != FLT_MAX && ds + w == dist[dst] && src != dst && (w > 0.f || src < dst);
    }
    __device__ __forceinline__ bool commit(int src, int dst, uint32_t, uint32_t, uint32_t) const {
        atomicMin(preds + dst, src);
        return false;
    }
};

// get_value_to_reduce of pr_functor_t (pr_functor.hxx:27-29)
struct FiniteValueFn {
    const float *values;
    __device__ __forceinline__ float operator()(int, int nbr, uint32_t) const {
        const float x = values[nbr];
        return isfinite(x) ? x : 0.0f;
    }
};

// The pull index array of a reduce and whether it is the remapped copy of b200_graph::hot_indices (b200_frontier.h).
static_assert(QSEG_HOT_MAX == B200_HOT_MAX, "operators.cuh and b200_frontier.h disagree");
bool use_hot_columns(b200_ctx *ctx, const b200_graph *g, int push) {
    return !push && ctx->adv_impl != B200_ADVANCE_LBS && g->hot_indices && g->hot_ids && g->hot_count > 0 &&
           g->hot_count <= B200_HOT_MAX && quad_aligned(g->hot_indices);
}
int ensure_hot_vals(b200_ctx *ctx) {
    if (!ctx->hot_vals) B200_CUDA(cudaMalloc(&ctx->hot_vals, sizeof(float) * B200_HOT_MAX));
    return B200_OK;
}

int check_counts(b200_ctx *ctx, int64_t len) {
    if (len < 0 || len >= (1ll << 32)) return B200_ERR_INVALID;
    return b200_ctx_reserve(ctx, len);
}

int run_compact_result(b200_ctx *ctx, int64_t *out_len) {
    B200_CUDA(read_counters(&ctx->ws));
    if (out_len) *out_len = (int64_t)ctx->ws.h_counters[B200_CNT_OUT];
    return ctx->ws.h_counters[B200_CNT_OVERFLOW] ? B200_ERR_OVERFLOW : B200_OK;
}

template <class Pred>
int compact_op(b200_ctx *ctx, Pred pred, int64_t len, int32_t *d_out, int64_t cap, int64_t *out_len) {
    if (len == 0) {
        if (out_len) *out_len = 0;
        return B200_OK;
    }
    B200_TRY(check_counts(ctx, len));
    b200_workspace *ws = &ctx->ws;
    B200_CUDA(cudaSetDevice(ws->device));
    B200_CUDA(reset_counters(ws));
    B200_CUDA(launch_compact(ws, pred, (uint32_t)len, d_out, (unsigned long long)cap, ws->d_counters + B200_CNT_OUT,
                             ws->d_counters + B200_CNT_OVERFLOW));
    return run_compact_result(ctx, out_len);
}

cudaEvent_t *level_events(b200_ctx *ctx) {
    if (!ctx->ev_level) {
        const int cnt = 3 * B200_MAX_LEVELS + 3;
        ctx->ev_level = new cudaEvent_t[cnt];
        for (int i = 0; i < cnt; ++i) cudaEventCreate(&ctx->ev_level[i]);
        ctx->ev_level_count = cnt;
    }
    return ctx->ev_level;
}

void finish_level_timing(b200_ctx *ctx, b200_stats *stats) {
    if (!stats || !stats->collect_timing) return;
    cudaEvent_t *ev = ctx->ev_level;
    const int L = stats->num_levels < B200_MAX_LEVELS ? stats->num_levels : B200_MAX_LEVELS;
    for (int l = 0; l < L; ++l) {
        cudaEventElapsedTime(&stats->level[l].advance_ms, ev[3 * l + 1], ev[3 * l + 2]);
        cudaEventElapsedTime(&stats->level[l].level_ms, ev[3 * l], ev[3 * l + 3]);
    }
}

}  // namespace

extern "C" {

// ============================================================== operators
int b200_advance_forward(b200_ctx *ctx, const b200_graph *g, const b200_problem *p, const int32_t *d_in,
                         int64_t in_len, int32_t *d_out, int64_t out_capacity, int iteration, int flags,
                         int64_t *out_len, int64_t *arcs) {
    if (!ctx || !g || !p || in_len < 0 || (in_len && !d_in)) return B200_ERR_INVALID;
    if (out_len) *out_len = 0;
    if (arcs) *arcs = 0;
    if (in_len == 0) return B200_OK;
    const bool no_out = flags & B200_ADV_NO_OUTPUT, raw = flags & B200_ADV_RAW_OUTPUT, idem = flags & B200_ADV_IDEMPOTENT;
    if (!no_out && !d_out) return B200_ERR_INVALID;
    B200_TRY(check_counts(ctx, in_len));
    b200_workspace *ws = &ctx->ws;
    B200_CUDA(cudaSetDevice(ws->device));
    B200_CUDA(reset_counters(ws));
    const unsigned long long cap = (unsigned long long)out_capacity;
    const bool quad = ctx->adv_impl != B200_ADVANCE_LBS && !raw && quad_aligned(g->col_indices) &&
                      (p->kind != B200_PROBLEM_SSSP || quad_aligned(p->weights));
    cudaError_t e = cudaErrorInvalidValue;
    if (quad) {
        B200_CUDA(launch_quad_scan(ws, d_in, (uint32_t)in_len, g->row_offsets));
        if (p->kind == B200_PROBLEM_BFS) {
            const QuadArgs a = make_quad_args(ws, d_in, (uint32_t)in_len, g->row_offsets, g->col_indices, nullptr);
            if (!p->visited_bitmap || (!idem && !p->labels)) return B200_ERR_INVALID;
            if (idem) {
                BfsIdempotentQ op{p->visited_bitmap};
                e = no_out ? launch_quad_advance<OUT_NONE, false>(ws, a, op, d_out, cap)
                           : launch_quad_advance<OUT_COMPACT, false>(ws, a, op, d_out, cap);
            } else {
                BfsPushQ op{p->visited_bitmap, p->labels, iteration + 1};
                e = no_out ? launch_quad_advance<OUT_NONE, false>(ws, a, op, d_out, cap)
                           : launch_quad_advance<OUT_COMPACT, false>(ws, a, op, d_out, cap);
            }
        } else if (p->kind == B200_PROBLEM_SSSP) {
            if (!p->dist || !p->weights || (!idem && !p->visited)) return B200_ERR_INVALID;
            const QuadArgs a = make_quad_args(ws, d_in, (uint32_t)in_len, g->row_offsets, g->col_indices, p->weights);
            SsspRelaxQ op{p->dist, p->preds, idem ? nullptr : p->visited, iteration};
            e = no_out ? launch_quad_advance<OUT_NONE, false>(ws, a, op, d_out, cap)
                       : launch_quad_advance<OUT_COMPACT, false>(ws, a, op, d_out, cap);
        } else {
            return B200_ERR_UNSUPPORTED;
        }
        B200_CUDA(e);
        B200_CUDA(read_counters(ws));
        if (arcs) *arcs = (int64_t)ws->h_counters[B200_CNT_ARCS];
        if (out_len) *out_len = no_out ? 0 : (int64_t)ws->h_counters[B200_CNT_OUT];
        return ws->h_counters[B200_CNT_OVERFLOW] ? B200_ERR_OVERFLOW : B200_OK;
    }
    B200_CUDA(launch_frontier_scan(ws, d_in, (uint32_t)in_len, g->row_offsets));
    const LbsArgs a = make_lbs_args(ws, d_in, (uint32_t)in_len, g->row_offsets, g->col_indices);
    if (p->kind == B200_PROBLEM_BFS) {
        if (idem) {
            if (!p->visited_bitmap) return B200_ERR_INVALID;
            BfsIdempotentOp op{p->visited_bitmap};
            e = no_out ? launch_lbs_advance<OUT_NONE, false>(ws, a, op, d_out, cap)
                       : raw ? launch_lbs_advance<OUT_RAW, false>(ws, a, op, d_out, cap)
                             : launch_lbs_advance<OUT_COMPACT, false>(ws, a, op, d_out, cap);
        } else {
            if (!p->visited_bitmap || !p->labels) return B200_ERR_INVALID;
            BfsPushOp op{p->visited_bitmap, p->labels, iteration + 1};
            e = no_out ? launch_lbs_advance<OUT_NONE, false>(ws, a, op, d_out, cap)
                       : raw ? launch_lbs_advance<OUT_RAW, false>(ws, a, op, d_out, cap)
                             : launch_lbs_advance<OUT_COMPACT, false>(ws, a, op, d_out, cap);
        }
    } else if (p->kind == B200_PROBLEM_SSSP) {
        if (!p->dist || !p->weights) return B200_ERR_INVALID;
        // idempotent => duplicates are kept in the output and b200_filter dedupes (config 3);
        // otherwise the visited stamp is applied inside the advance.
        SsspRelaxOp op{p->dist, p->weights, p->preds, idem ? nullptr : p->visited, iteration};
        if (!idem && !p->visited) return B200_ERR_INVALID;
        e = no_out ? launch_lbs_advance<OUT_NONE, false>(ws, a, op, d_out, cap)
                   : raw ? launch_lbs_advance<OUT_RAW, false>(ws, a, op, d_out, cap)
                         : launch_lbs_advance<OUT_COMPACT, false>(ws, a, op, d_out, cap);
    } else {
        return B200_ERR_UNSUPPORTED;
    }
    B200_CUDA(e);
    B200_CUDA(read_counters(ws));
    if (arcs) *arcs = (int64_t)ws->h_counters[B200_CNT_TOTAL];
    if (out_len) *out_len = no_out ? 0 : (int64_t)ws->h_counters[B200_CNT_OUT];
    return ws->h_counters[B200_CNT_OVERFLOW] ? B200_ERR_OVERFLOW : B200_OK;
}

int b200_filter(b200_ctx *ctx, const b200_graph *g, const b200_problem *p, const int32_t *d_in, int64_t in_len,
                int32_t *d_out, int64_t out_capacity, int iteration, int64_t *out_len) {
    if (!ctx || !p || in_len < 0 || (in_len && (!d_in || !d_out))) return B200_ERR_INVALID;
    switch (p->kind) {
        case B200_PROBLEM_BFS:
            return compact_op(ctx, BfsFilterPred{d_in}, in_len, d_out, out_capacity, out_len);
        case B200_PROBLEM_SSSP:
            if (!p->visited) return B200_ERR_INVALID;
            return compact_op(ctx, SsspFilterPred{d_in, p->visited, iteration}, in_len, d_out, out_capacity, out_len);
        case B200_PROBLEM_PR:
            if (!p->current_ranks || !p->reduced_ranks || (!p->degrees && !g)) return B200_ERR_INVALID;
            return compact_op(ctx, PrFilterPred{d_in, p->current_ranks, p->reduced_ranks, p->degrees, g ? g->row_offsets : nullptr},
                              in_len, d_out, out_capacity, out_len);
        default:
            return B200_ERR_UNSUPPORTED;
    }
}

int b200_uniquify(b200_ctx *ctx, const b200_graph *, const b200_problem *p, uint32_t *d_visited_bitmap,
                  const int32_t *d_in, int64_t in_len, int32_t *d_out, int64_t out_capacity, int iteration,
                  int64_t *out_len) {
    if (!ctx || !d_visited_bitmap || in_len < 0 || (in_len && (!d_in || !d_out))) return B200_ERR_INVALID;
    int *labels = (p && p->kind == B200_PROBLEM_BFS) ? p->labels : nullptr;
    return compact_op(ctx, UniquifyPred{d_in, d_visited_bitmap, labels, iteration + 1}, in_len, d_out, out_capacity, out_len);
}

int b200_sparse_to_dense(b200_ctx *ctx, int64_t n, const int32_t *d_sparse, int64_t len, uint32_t *d_bitmap) {
    if (!ctx || n < 0 || len < 0 || !d_bitmap || (len && !d_sparse)) return B200_ERR_INVALID;
    cudaStream_t st = ws_stream(&ctx->ws);
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    B200_CUDA(cudaMemsetAsync(d_bitmap, 0, sizeof(uint32_t) * (size_t)((n + 31) / 32), st));
    if (len) {
        sparse_to_bitmap_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(d_sparse, (uint32_t)len, d_bitmap);
        ctx->ws.launches++;
        B200_CUDA(cudaGetLastError());
    }
    B200_CUDA(cudaStreamSynchronize(st));
    return B200_OK;
}

int b200_dense_to_sparse(b200_ctx *ctx, int64_t n, const uint32_t *d_bitmap, int32_t *d_sparse, int64_t capacity,
                         int64_t *out_len) {
    if (!ctx || !d_bitmap || !d_sparse) return B200_ERR_INVALID;
    if (out_len) *out_len = 0;
    if (n <= 0) return n == 0 ? B200_OK : B200_ERR_INVALID;
    B200_TRY(check_counts(ctx, n));
    b200_workspace *ws = &ctx->ws;
    B200_CUDA(cudaSetDevice(ws->device));
    B200_CUDA(reset_counters(ws));
    // word-wise (tile_scan.cuh bitmap_list_kernel); bits past n in the last word must be clear
    B200_CUDA(launch_bitmap_list(ws, BitmapWords{d_bitmap}, IdentityItem{}, (uint32_t)n, d_sparse, (unsigned long long)capacity,
                                 ws->d_counters + B200_CNT_OUT, ws->d_counters + B200_CNT_OVERFLOW));
    B200_CUDA(read_counters(ws));
    if (ws->h_counters[B200_CNT_OVERFLOW]) return B200_ERR_OVERFLOW;
    if (out_len) *out_len = (int64_t)ws->h_counters[B200_CNT_OUT];
    return B200_OK;
}

int b200_gen_unvisited(b200_ctx *ctx, const b200_problem *p, int64_t n, int32_t *d_unvisited, int64_t capacity,
                       int64_t *out_len) {
    if (!ctx || !p || p->kind != B200_PROBLEM_BFS || !p->labels || !d_unvisited) return B200_ERR_INVALID;
    return compact_op(ctx, UnvisitedPred{p->labels}, n, d_unvisited, capacity, out_len);
}

int b200_advance_backward(b200_ctx *ctx, const b200_graph *g, const b200_problem *p, const uint32_t *d_frontier_bitmap,
                          uint32_t *d_next_bitmap, int iteration, int64_t *discovered, int64_t *arcs_inspected) {
    if (!ctx || !g || !p || p->kind != B200_PROBLEM_BFS || !p->labels || !p->visited_bitmap || !d_frontier_bitmap ||
        !d_next_bitmap)
        return B200_ERR_INVALID;
    b200_workspace *ws = &ctx->ws;
    B200_CUDA(cudaSetDevice(ws->device));
    B200_CUDA(reset_counters(ws));
    const uint32_t *off = g->col_offsets ? g->col_offsets : g->row_offsets;
    const int32_t *idx = g->row_indices ? g->row_indices : g->col_indices;
    bfs_pull_kernel<256><<<ws->num_sms * 8, 256, 0, ws_stream(ws)>>>((uint32_t)g->n, off, idx, d_frontier_bitmap, d_next_bitmap,
                                                                      p->visited_bitmap, p->labels, iteration + 1, ws->d_counters, Partition{0, 0, (uint32_t)g->n});
    ws->launches++;
    B200_CUDA(cudaGetLastError());
    B200_CUDA(read_counters(ws));
    if (discovered) *discovered = (int64_t)ws->h_counters[B200_CNT_OUT];
    if (arcs_inspected) *arcs_inspected = (int64_t)ws->h_counters[B200_CNT_ARCS];
    return B200_OK;
}

int b200_neighborhood_reduce_f32(b200_ctx *ctx, const b200_graph *g, const int32_t *d_in, int64_t in_len,
                                 const float *d_values, float *d_reduced, float identity, int op, int push,
                                 int scatter, int64_t *arcs) {
    if (!ctx || !g || in_len < 0 || (in_len && (!d_in || !d_values || !d_reduced))) return B200_ERR_INVALID;
    if (arcs) *arcs = 0;
    if (in_len == 0) return B200_OK;
    B200_TRY(check_counts(ctx, in_len));
    b200_workspace *ws = &ctx->ws;
    B200_CUDA(cudaSetDevice(ws->device));
    const uint32_t *off = push ? g->row_offsets : (g->col_offsets ? g->col_offsets : g->row_offsets);
    const int32_t *idx = push ? g->col_indices : (g->row_indices ? g->row_indices : g->col_indices);
    B200_CUDA(reset_counters(ws));
    const float neutral = op == B200_OP_PLUS ? PlusF32::neutral() : (op == B200_OP_MIN ? MinF32::neutral() : MaxF32::neutral());
    FiniteValueFn vf{d_values};
    cudaError_t e;
    const bool quad = ctx->adv_impl != B200_ADVANCE_LBS && quad_aligned(idx);
    if (quad && use_hot_columns(ctx, g, push)) {
        // hot columns: the remapped index copy, hot values served from shared memory (quad_segreduce.cuh)
        B200_TRY(ensure_hot_vals(ctx));
        B200_CUDA(launch_neighborhood_quad_scan<float>(ws, d_in, (uint32_t)in_len, off, d_reduced, identity, neutral, scatter));
        const QuadArgs a = make_quad_args(ws, d_in, (uint32_t)in_len, off, g->hot_indices, nullptr);
        const uint32_t hc = (uint32_t)g->hot_count;
        if (op == B200_OP_PLUS) e = launch_quad_segreduce_hot<float, PlusF32>(ws, a, vf, d_reduced, scatter, g->hot_ids, hc, ctx->hot_vals);
        else if (op == B200_OP_MIN) e = launch_quad_segreduce_hot<float, MinF32>(ws, a, vf, d_reduced, scatter, g->hot_ids, hc, ctx->hot_vals);
        else if (op == B200_OP_MAX) e = launch_quad_segreduce_hot<float, MaxF32>(ws, a, vf, d_reduced, scatter, g->hot_ids, hc, ctx->hot_vals);
        else return B200_ERR_INVALID;
    } else if (quad) {
        B200_CUDA(launch_neighborhood_quad_scan<float>(ws, d_in, (uint32_t)in_len, off, d_reduced, identity, neutral, scatter));
        const QuadArgs a = make_quad_args(ws, d_in, (uint32_t)in_len, off, idx, nullptr);
        if (op == B200_OP_PLUS) e = launch_quad_segreduce<float, PlusF32>(ws, a, vf, d_reduced, scatter);
        else if (op == B200_OP_MIN) e = launch_quad_segreduce<float, MinF32>(ws, a, vf, d_reduced, scatter);
        else if (op == B200_OP_MAX) e = launch_quad_segreduce<float, MaxF32>(ws, a, vf, d_reduced, scatter);
        else return B200_ERR_INVALID;
    } else {
        NeighborhoodDegree<float> deg{d_in, off, d_reduced, identity, neutral, scatter};
        B200_CUDA(launch_scan(ws, deg, (uint32_t)in_len, ws->d_scanned, ws->d_counters + B200_CNT_TOTAL));
        const LbsArgs a = make_lbs_args(ws, d_in, (uint32_t)in_len, off, idx);
        if (op == B200_OP_PLUS) e = launch_lbs_segreduce<float, PlusF32>(ws, a, vf, d_reduced, scatter);
        else if (op == B200_OP_MIN) e = launch_lbs_segreduce<float, MinF32>(ws, a, vf, d_reduced, scatter);
        else if (op == B200_OP_MAX) e = launch_lbs_segreduce<float, MaxF32>(ws, a, vf, d_reduced, scatter);
        else return B200_ERR_INVALID;
    }
    B200_CUDA(e);
    B200_CUDA(read_counters(ws));
    if (arcs) *arcs = (int64_t)ws->h_counters[quad ? B200_CNT_ARCS : B200_CNT_TOTAL];
    return B200_OK;
}

// ============================================================== primitives
int b200_bfs_run(b200_ctx *ctx, const b200_graph *g, int32_t src, int mode, float alpha, float beta,
                 int32_t *d_labels, b200_stats *stats) {
    if (!ctx || !g || !d_labels || g->n < 1 || src < 0 || src >= g->n || g->n > (1ll << 31)) return B200_ERR_INVALID;
    if (mode < B200_BFS_PUSH || mode > B200_BFS_BEAMER) return B200_ERR_INVALID;
    const int64_t n = g->n;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    B200_TRY(ensure_traversal_scratch(ctx, n));
    b200_workspace *ws = &ctx->ws;
    cudaStream_t st = ws_stream(ws);
    const size_t words = (size_t)((n + 31) / 32);
    const bool timing = stats && stats->collect_timing;
    const bool quad = ctx->adv_impl != B200_ADVANCE_LBS && quad_aligned(g->col_indices);
    if (alpha <= 0.f) alpha = 15.f;
    if (beta <= 0.f) beta = 18.f;
    if (quad && !timing && ctx->loop_impl == B200_LOOP_GRAPH) {
        const int gs = bfs_run_graph(ctx, g, src, mode, alpha, beta, d_labels, stats);
        if (gs != B200_ERR_UNSUPPORTED) return gs;   // unsupported: the graph could not be built, run the host loop
    }
    if (stats) stats->level_loop = B200_LOOP_HOST;
    {   // experiment knob (profiles/README.md): persisting-L2 access-policy window over the visited bitmap, host-driven loop
        const char *pin = getenv("B200_L2_PIN_VISITED");
        if (pin && pin[0] == '1' && !ctx->l2_window_set) B200_TRY(b200_ctx_l2_pin(ctx, ctx->bm_visited, (int64_t)(words * sizeof(uint32_t))));
    }
    cudaEvent_t *ev = timing ? level_events(ctx) : nullptr;
    const int64_t launches0 = ws->launches;
    const uint32_t *pull_off = g->col_offsets ? g->col_offsets : g->row_offsets;
    const int32_t *pull_idx = g->row_indices ? g->row_indices : g->col_indices;

    B200_CUDA(cudaEventRecord(ctx->ev_run[0], st));
    // bfs_problem_t ctor (bfs_problem.hxx:38-42) + init_frontier (bfs_enactor.hxx:34-38)
    B200_CUDA(cudaMemsetAsync(d_labels, 0xFF, sizeof(int32_t) * (size_t)n, st));
    B200_CUDA(cudaMemsetAsync(ctx->bm_visited, 0, sizeof(uint32_t) * words, st));
    bfs_init_kernel<<<1, 1, 0, st>>>(d_labels, ctx->bm_visited, ctx->frontier[0], src);
    ws->launches++;
    B200_CUDA(cudaGetLastError());

    int sel = 0, bsel = 0, level = 0, wsel = 0;
    const bool work_create = quad && ctx->adv_impl == B200_ADVANCE_QUAD && n <= WORK_CREATE_MAX_N;   // (else: scan before every level)
    bool have_scan = false;
    int64_t next_quads = 0;
    int64_t flen = 1, unvisited = n - 1, reached = 1, total_arcs = 0;
    int64_t m_unexplored = g->m;
    bool pull = false;

    for (;;) {
        b200_level_stat *ls = (stats && level < B200_MAX_LEVELS) ? &stats->level[level] : nullptr;
        const bool tl = timing && level < B200_MAX_LEVELS;
        if (tl && level == 0) B200_CUDA(cudaEventRecord(ev[0], st));   // later levels start at the previous level's end marker
        B200_CUDA(reset_counters(ws));
        int64_t found, arcs, next_deg;
        if (!pull) {
            const bool deg = mode == B200_BFS_BEAMER;
            int32_t *next = ctx->frontier[sel ^ 1];
            if (quad) {
                if (!have_scan) {
                    B200_CUDA(launch_quad_scan(ws, ctx->frontier[sel], (uint32_t)flen, g->row_offsets));   // into pair 0
                    wsel = 0;
                } else {   // the previous advance created this level's scan and row bounds (pair wsel)
                    set_counter_kernel<<<1, 1, 0, st>>>(ws->d_counters + B200_CNT_TOTAL, (unsigned long long)next_quads);
                    ws->launches++;
                }
                QuadArgs a = make_quad_args(ws, ctx->frontier[sel], (uint32_t)flen, g->row_offsets, g->col_indices, nullptr);
                if (work_create) {
                    uint32_t *sc[2] = {ws->d_scanned, ws->d_scanned2};
                    uint2 *rw[2] = {reinterpret_cast<uint2 *>(ws->d_rows), reinterpret_cast<uint2 *>(ws->d_rows2)};
                    a.scanned = sc[wsel];
                    a.rows = rw[wsel];
                    a.scanned_next = sc[wsel ^ 1];
                    a.rows_next = rw[wsel ^ 1];
                }
                BfsPushQ op{ctx->bm_visited, d_labels, level + 1};
                if (tl) B200_CUDA(cudaEventRecord(ev[3 * level + 1], st));
                if (deg) B200_CUDA((launch_quad_advance<OUT_COMPACT, true>(ws, a, op, next, (unsigned long long)n)));
                else B200_CUDA((launch_quad_advance<OUT_COMPACT, false>(ws, a, op, next, (unsigned long long)n)));
            } else {
                B200_CUDA(launch_frontier_scan(ws, ctx->frontier[sel], (uint32_t)flen, g->row_offsets));
                const LbsArgs a = make_lbs_args(ws, ctx->frontier[sel], (uint32_t)flen, g->row_offsets, g->col_indices);
                BfsPushOp op{ctx->bm_visited, d_labels, level + 1};
                if (tl) B200_CUDA(cudaEventRecord(ev[3 * level + 1], st));
                if (deg) B200_CUDA((launch_lbs_advance<OUT_COMPACT, true>(ws, a, op, next, (unsigned long long)n)));
                else B200_CUDA((launch_lbs_advance<OUT_COMPACT, false>(ws, a, op, next, (unsigned long long)n)));
            }
            if (tl) B200_CUDA(cudaEventRecord(ev[3 * level + 2], st));
            B200_CUDA(read_counters(ws));
            if (ws->h_counters[B200_CNT_OVERFLOW]) return B200_ERR_OVERFLOW;
            found = (int64_t)ws->h_counters[B200_CNT_OUT];
            if (work_create) {   // (vertices emitted << 32) | quads created
                next_quads = found & 0xffffffffll;
                found = (int64_t)((unsigned long long)found >> 32);
                have_scan = true;
                wsel ^= 1;
            }
            arcs = (int64_t)ws->h_counters[quad ? B200_CNT_ARCS : B200_CNT_TOTAL];
            next_deg = (int64_t)ws->h_counters[B200_CNT_AUX];
        } else {
            have_scan = false;   // the list that follows a pull phase comes from the bitmap: it needs a scan
            if (tl) B200_CUDA(cudaEventRecord(ev[3 * level + 1], st));
            bfs_pull_kernel<256><<<ws->num_sms * 8, 256, 0, st>>>((uint32_t)n, pull_off, pull_idx, ctx->bm_frontier[bsel],
                                                                  ctx->bm_frontier[bsel ^ 1], ctx->bm_visited, d_labels,
                                                                  level + 1, ws->d_counters, Partition{0, 0, (uint32_t)n},
                                                                  g->first_in_neighbor);
            ws->launches++;
            B200_CUDA(cudaGetLastError());
            if (tl) B200_CUDA(cudaEventRecord(ev[3 * level + 2], st));
            B200_CUDA(read_counters(ws));
            found = (int64_t)ws->h_counters[B200_CNT_OUT];
            arcs = (int64_t)ws->h_counters[B200_CNT_ARCS];
            next_deg = (int64_t)ws->h_counters[B200_CNT_AUX];
        }
        if (ls) {
            ls->direction = pull ? 1 : 0;
            ls->frontier_len = pull ? unvisited : flen;
            ls->arcs = arcs;
            ls->discovered = found;
        }
        total_arcs += arcs;
        ++level;
        if (tl) B200_CUDA(cudaEventRecord(ev[3 * level], st));   // closes level-1; reused as next level's start marker
        if (found == 0) break;
        reached += found;
        unvisited -= found;
        const bool was_pull = pull;
        if (!pull) {
            m_unexplored -= arcs;
            sel ^= 1;
            if (mode == B200_BFS_REF_ALPHA) {
                // bfs_enactor.hxx:68: (float)num_unvisited < (float)frontier_length * threshold
                if ((float)unvisited < (float)found * alpha) pull = true;
            } else if (mode == B200_BFS_BEAMER) {
                if ((double)next_deg > (double)m_unexplored / alpha && found > flen) pull = true;
            }
        } else {
            bsel ^= 1;
            if (mode == B200_BFS_BEAMER && (double)found < (double)n / beta && found < flen) pull = false;
        }
        flen = found;
        if (pull && !was_pull) {
            // vertices without in-arcs count as visited from here on: pull levels skip them (engine.cuh)
            B200_CUDA(launch_or_no_in_arc(ws, pull_off, n, g->no_in_arc_bitmap, ctx->bm_visited));
            // frontier list -> bitmap (sparse_to_dense_kernel, advance.hxx:69-84)
            B200_CUDA(cudaMemsetAsync(ctx->bm_frontier[bsel], 0, sizeof(uint32_t) * words, st));
            sparse_to_bitmap_kernel<<<(unsigned)((flen + 255) / 256), 256, 0, st>>>(ctx->frontier[sel], (uint32_t)flen,
                                                                                   ctx->bm_frontier[bsel]);
            ws->launches++;
            B200_CUDA(cudaGetLastError());
        } else if (!pull && was_pull) {
            // bitmap -> frontier list for the push levels that finish the traversal
            B200_CUDA(reset_counters(ws));
            B200_CUDA(launch_bitmap_list(ws, BitmapWords{ctx->bm_frontier[bsel]}, IdentityItem{}, (uint32_t)n, ctx->frontier[sel],
                                         (unsigned long long)n, ws->d_counters + B200_CNT_OUT, ws->d_counters + B200_CNT_OVERFLOW));
        }
    }
    B200_CUDA(cudaEventRecord(ctx->ev_run[1], st));
    B200_CUDA(cudaEventSynchronize(ctx->ev_run[1]));
    if (stats) {
        stats->num_levels = level;
        stats->reached = reached;
        stats->total_arcs = total_arcs;
        stats->launches = ws->launches - launches0;
        B200_CUDA(cudaEventElapsedTime(&stats->device_ms, ctx->ev_run[0], ctx->ev_run[1]));
        finish_level_timing(ctx, stats);
    }
    return B200_OK;
}

// First bucket width of the near-far loop (near_far.cuh): what b200_ctx_set_sssp_delta chose, else 3.5 * mean weight
// / mean degree from a strided sample of <= 65536 weights, computed once per (weights array, n, m).
static int sssp_delta0(b200_ctx *ctx, const b200_graph *g, float *out) {
    if (ctx->sssp_delta > 0.f) {
        *out = ctx->sssp_delta;
        return B200_OK;
    }
    if (ctx->nf_key_w != g->col_values || ctx->nf_key_n != g->n || ctx->nf_key_m != g->m) {
        float delta = INFINITY;
        if (g->m > 0) {
            cudaStream_t st = ws_stream(&ctx->ws);
            const unsigned long long m = (unsigned long long)g->m;
            const unsigned long long step = m > 65536ull ? m / 65536ull : 1ull;
            const unsigned int samples = (unsigned int)((m + step - 1) / step);
            double sum = 0.0;
            B200_CUDA(cudaMemsetAsync(ctx->d_nf_sum, 0, sizeof(double), st));
            nf_weight_sample_kernel<<<(samples + 255) / 256, 256, 0, st>>>(g->col_values, m, step, samples, ctx->d_nf_sum);
            B200_CUDA(cudaGetLastError());
            B200_CUDA(cudaMemcpyAsync(&sum, ctx->d_nf_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            const double d = 3.5 * (sum / (double)samples) * (double)g->n / (double)g->m;
            if (d > 0.0 && d < 1e30) delta = (float)d;   // all-zero (or non-finite) weights: one bucket = Bellman-Ford
        }
        ctx->nf_key_w = g->col_values;
        ctx->nf_key_n = g->n;
        ctx->nf_key_m = g->m;
        ctx->nf_auto_delta = delta;
    }
    *out = ctx->nf_auto_delta;
    return B200_OK;
}

int b200_sssp_run(b200_ctx *ctx, const b200_graph *g, int32_t src, float *d_dist, int32_t *d_preds, b200_stats *stats) {
    if (!ctx || !g || !d_dist || !g->col_values || g->n < 1 || src < 0 || src >= g->n || g->n > (1ll << 31))
        return B200_ERR_INVALID;
    const int64_t n = g->n;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    B200_TRY(ensure_traversal_scratch(ctx, n));
    b200_workspace *ws = &ctx->ws;
    cudaStream_t st = ws_stream(ws);
    const bool timing = stats && stats->collect_timing;
    cudaEvent_t *ev = timing ? level_events(ctx) : nullptr;
    const int64_t launches0 = ws->launches;

    const bool quad = ctx->adv_impl != B200_ADVANCE_LBS && quad_aligned(g->col_indices) && quad_aligned(g->col_values);
    const bool work_create = quad && ctx->adv_impl == B200_ADVANCE_QUAD && n <= WORK_CREATE_MAX_N;
    bool graph_done = false, have_scan = false;
    int64_t next_quads = 0;
    int wsel = 0;
    int sel = 0, it = 0;
    int64_t flen = 1, total_arcs = 0, bucket_arcs = 0;
    // near-far ordering of the frontier iterations (near_far.cuh); the arc-wise kernel keeps the reference's order
    float delta0 = INFINITY;
    if (quad) B200_TRY(sssp_delta0(ctx, g, &delta0));
    const NearFar *nf = (quad && delta0 < INFINITY) ? ctx->d_nf : nullptr;
    if (quad && !timing && ctx->loop_impl == B200_LOOP_GRAPH) {
        // the frontier iterations as one CUDA graph (level_loop.cu); the predecessor pass follows below
        b200_stats *gst = stats ? stats : new (std::nothrow) b200_stats;
        if (!gst) return B200_ERR_NOMEM;
        const int gs = sssp_run_graph(ctx, g, src, d_dist, delta0, gst);
        if (gs == B200_OK) {
            graph_done = true;
            it = gst->num_levels;
            total_arcs = gst->total_arcs;
        }
        if (!stats) delete gst;
        if (gs != B200_OK && gs != B200_ERR_UNSUPPORTED) {
            return gs;
        }
    }
    if (!graph_done) {
    if (stats) stats->level_loop = B200_LOOP_HOST;
    B200_CUDA(cudaEventRecord(ctx->ev_run[0], st));
    // sssp_problem_t ctor (sssp_problem.hxx:44-49) + init_frontier (sssp_enactor.hxx:33-37)
    sssp_init_kernel<<<ws->num_sms * 4, 256, 0, st>>>(d_dist, d_preds, ctx->stamp, (unsigned long long)n);
    sssp_seed_kernel<<<1, 1, 0, st>>>(d_dist, ctx->frontier[0], src);
    ws->launches += 2;
    if (nf) {
        nf_init_kernel<<<1, 1, 0, st>>>(ctx->d_nf, delta0, (long long)g->m);
        ws->launches++;
    }
    B200_CUDA(cudaGetLastError());
    for (;;) {
        b200_level_stat *ls = (stats && it < B200_MAX_LEVELS) ? &stats->level[it] : nullptr;
        const bool tl = timing && it < B200_MAX_LEVELS;
        if (tl && it == 0) B200_CUDA(cudaEventRecord(ev[0], st));
        B200_CUDA(reset_counters(ws));
        if (quad) {
            if (!have_scan) {
                B200_CUDA(launch_quad_scan(ws, ctx->frontier[sel], (uint32_t)flen, g->row_offsets));
                wsel = 0;
            } else {
                set_counter_kernel<<<1, 1, 0, st>>>(ws->d_counters + B200_CNT_TOTAL, (unsigned long long)next_quads);
                ws->launches++;
            }
            QuadArgs a = make_quad_args(ws, ctx->frontier[sel], (uint32_t)flen, g->row_offsets, g->col_indices, g->col_values);
            if (work_create) {
                uint32_t *sc[2] = {ws->d_scanned, ws->d_scanned2};
                uint2 *rw[2] = {reinterpret_cast<uint2 *>(ws->d_rows), reinterpret_cast<uint2 *>(ws->d_rows2)};
                a.scanned = sc[wsel];
                a.rows = rw[wsel];
                a.scanned_next = sc[wsel ^ 1];
                a.rows_next = rw[wsel ^ 1];
            }
            SsspRelaxQ op{d_dist, nullptr, ctx->stamp, it, nf};   // preds: one exact pass at the end
            if (tl) B200_CUDA(cudaEventRecord(ev[3 * it + 1], st));
            B200_CUDA((launch_quad_advance<OUT_COMPACT, false>(ws, a, op, ctx->frontier[sel ^ 1], (unsigned long long)n)));
        } else {
            B200_CUDA(launch_frontier_scan(ws, ctx->frontier[sel], (uint32_t)flen, g->row_offsets));
            const LbsArgs a = make_lbs_args(ws, ctx->frontier[sel], (uint32_t)flen, g->row_offsets, g->col_indices);
            SsspRelaxOp op{d_dist, g->col_values, nullptr, ctx->stamp, it};   // preds: one exact pass at the end
            if (tl) B200_CUDA(cudaEventRecord(ev[3 * it + 1], st));
            B200_CUDA((launch_lbs_advance<OUT_COMPACT, false>(ws, a, op, ctx->frontier[sel ^ 1], (unsigned long long)n)));
        }
        if (tl) B200_CUDA(cudaEventRecord(ev[3 * it + 2], st));
        B200_CUDA(read_counters(ws));
        if (ws->h_counters[B200_CNT_OVERFLOW]) return B200_ERR_OVERFLOW;
        int64_t found = (int64_t)ws->h_counters[B200_CNT_OUT];
        if (work_create) {
            next_quads = found & 0xffffffffll;
            found = (int64_t)((unsigned long long)found >> 32);
            have_scan = true;
            wsel ^= 1;
        }
        const int64_t arcs = (int64_t)ws->h_counters[quad ? B200_CNT_ARCS : B200_CNT_TOTAL];
        if (ls) {
            ls->direction = 0;
            ls->frontier_len = flen;
            ls->arcs = arcs;
            ls->discovered = found;
        }
        total_arcs += arcs;
        bucket_arcs += arcs;
        ++it;
        if (found == 0 && nf) {
            // the near frontier ran dry: every distance below the cutoff is final; open the next bucket
            NearFar h_nf;
            const NearFarHostHook hook{ctx->frontier[sel ^ 1], (long long)bucket_arcs};
            const unsigned grid = nf_grid(ws->num_sms, n);
            sssp_pending_min_kernel<<<grid, NF_NT, 0, st>>>(d_dist, (uint32_t)n, ctx->d_nf, hook);
            sssp_take_kernel<<<grid, NF_NT, 0, st>>>(d_dist, (uint32_t)n, ctx->d_nf, hook);
            ws->launches += 2;
            B200_CUDA(cudaGetLastError());
            B200_CUDA(cudaMemcpyAsync(&h_nf, ctx->d_nf, sizeof(NearFar), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            bucket_arcs = 0;
            have_scan = false;           // a list nobody has scanned
            if (!h_nf.done) found = (int64_t)h_nf.taken;
        }
        if (tl) B200_CUDA(cudaEventRecord(ev[3 * it], st));
        if (found == 0) break;
        sel ^= 1;
        flen = found;
    }
    }   // !graph_done
    if (d_preds) {
        iota_fill_kernel<<<ws->num_sms * 4, 256, 0, st>>>(ctx->frontier[0], d_preds, (unsigned long long)n);
        ws->launches++;
        B200_CUDA(cudaGetLastError());
        B200_CUDA(reset_counters(ws));
        if (quad) {
            B200_CUDA(launch_quad_scan(ws, ctx->frontier[0], (uint32_t)n, g->row_offsets));
            const QuadArgs a = make_quad_args(ws, ctx->frontier[0], (uint32_t)n, g->row_offsets, g->col_indices, g->col_values);
            B200_CUDA((launch_quad_advance<OUT_NONE, false>(ws, a, SsspPredQ{d_dist, d_preds}, nullptr, 0ull)));
        } else {
            B200_CUDA(launch_frontier_scan(ws, ctx->frontier[0], (uint32_t)n, g->row_offsets));
            const LbsArgs a = make_lbs_args(ws, ctx->frontier[0], (uint32_t)n, g->row_offsets, g->col_indices);
            B200_CUDA((launch_lbs_advance<OUT_NONE, false>(ws, a, SsspPredOp{d_dist, g->col_values, d_preds}, nullptr, 0ull)));
        }
        preds_fixup_kernel<<<ws->num_sms * 4, 256, 0, st>>>(d_preds, (unsigned long long)n, src);
        ws->launches++;
        B200_CUDA(cudaGetLastError());
    }
    B200_CUDA(cudaEventRecord(ctx->ev_run[1], st));
    B200_CUDA(cudaEventSynchronize(ctx->ev_run[1]));
    if (stats) {
        stats->num_levels = it;
        stats->reached = -1;
        stats->total_arcs = total_arcs;
        stats->launches = ws->launches - launches0;
        B200_CUDA(cudaEventElapsedTime(&stats->device_ms, ctx->ev_run[0], ctx->ev_run[1]));
        finish_level_timing(ctx, stats);
    }
    return B200_OK;
}

int b200_pr_run(b200_ctx *ctx, const b200_graph *g, int max_iter, int scatter, float *d_current, float *d_reduced,
                int64_t *h_frontier_lens, int *iterations, b200_stats *stats) {
    if (!ctx || !g || !d_current || !d_reduced || g->n < 1 || g->n > (1ll << 31)) return B200_ERR_INVALID;
    const int64_t n = g->n;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    B200_TRY(ensure_traversal_scratch(ctx, n));
    b200_workspace *ws = &ctx->ws;
    cudaStream_t st = ws_stream(ws);
    const bool timing = stats && stats->collect_timing;
    cudaEvent_t *ev = timing ? level_events(ctx) : nullptr;
    const int64_t launches0 = ws->launches;
    const uint32_t *off = g->col_offsets ? g->col_offsets : g->row_offsets;   // push=false (pr_enactor.hxx:53)
    const int32_t *idx = g->row_indices ? g->row_indices : g->col_indices;

    B200_CUDA(cudaEventRecord(ctx->ev_run[0], st));
    pr_init_kernel<<<ws->num_sms * 4, 256, 0, st>>>(d_current, d_reduced, ctx->frontier[0], (unsigned long long)n);
    ws->launches++;
    B200_CUDA(cudaGetLastError());
    int sel = 0, it = 0;
    int64_t flen = n, total_arcs = 0;
    const bool quad = ctx->adv_impl != B200_ADVANCE_LBS && quad_aligned(idx);
    const bool hot = quad && use_hot_columns(ctx, g, 0);
    if (hot) B200_TRY(ensure_hot_vals(ctx));
    while (flen > 0 && it < max_iter) {
        b200_level_stat *ls = (stats && it < B200_MAX_LEVELS) ? &stats->level[it] : nullptr;
        const bool tl = timing && it < B200_MAX_LEVELS;
        if (tl && it == 0) B200_CUDA(cudaEventRecord(ev[0], st));
        B200_CUDA(reset_counters(ws));
        if (quad) {
            B200_CUDA(launch_neighborhood_quad_scan<float>(ws, ctx->frontier[sel], (uint32_t)flen, off, d_reduced, 0.0f, 0.0f, scatter));
            const QuadArgs a = make_quad_args(ws, ctx->frontier[sel], (uint32_t)flen, off, hot ? g->hot_indices : idx, nullptr);
            if (tl) B200_CUDA(cudaEventRecord(ev[3 * it + 1], st));
            if (hot)
                B200_CUDA((launch_quad_segreduce_hot<float, PlusF32>(ws, a, FiniteValueFn{d_current}, d_reduced, scatter, g->hot_ids,
                                                                     (uint32_t)g->hot_count, ctx->hot_vals)));
            else
                B200_CUDA((launch_quad_segreduce<float, PlusF32>(ws, a, FiniteValueFn{d_current}, d_reduced, scatter)));
        } else {
            NeighborhoodDegree<float> deg{ctx->frontier[sel], off, d_reduced, 0.0f, 0.0f, scatter};
            B200_CUDA(launch_scan(ws, deg, (uint32_t)flen, ws->d_scanned, ws->d_counters + B200_CNT_TOTAL));
            const LbsArgs a = make_lbs_args(ws, ctx->frontier[sel], (uint32_t)flen, off, idx);
            if (tl) B200_CUDA(cudaEventRecord(ev[3 * it + 1], st));
            B200_CUDA((launch_lbs_segreduce<float, PlusF32>(ws, a, FiniteValueFn{d_current}, d_reduced, scatter)));
        }
        if (tl) B200_CUDA(cudaEventRecord(ev[3 * it + 2], st));
        B200_CUDA(launch_compact(ws, PrFilterPred{ctx->frontier[sel], d_current, d_reduced, nullptr, g->row_offsets},
                                 (uint32_t)flen, ctx->frontier[sel ^ 1], (unsigned long long)n,
                                 ws->d_counters + B200_CNT_OUT, ws->d_counters + B200_CNT_OVERFLOW));
        B200_CUDA(read_counters(ws));
        const int64_t out = (int64_t)ws->h_counters[B200_CNT_OUT];
        const int64_t arcs = (int64_t)ws->h_counters[quad ? B200_CNT_ARCS : B200_CNT_TOTAL];
        if (ls) {
            ls->direction = 1;
            ls->frontier_len = flen;
            ls->arcs = arcs;
            ls->discovered = out;
        }
        if (h_frontier_lens) h_frontier_lens[it] = out;
        total_arcs += arcs;
        ++it;
        if (tl) B200_CUDA(cudaEventRecord(ev[3 * it], st));
        sel ^= 1;
        flen = out;
    }
    B200_CUDA(cudaEventRecord(ctx->ev_run[1], st));
    B200_CUDA(cudaEventSynchronize(ctx->ev_run[1]));
    if (iterations) *iterations = it;
    if (stats) {
        stats->num_levels = it;
        stats->reached = n;
        stats->total_arcs = total_arcs;
        stats->launches = ws->launches - launches0;
        B200_CUDA(cudaEventElapsedTime(&stats->device_ms, ctx->ev_run[0], ctx->ev_run[1]));
        finish_level_timing(ctx, stats);
    }
    return B200_OK;
}

// ============================================================== host-buffer entry points
int b200_host_graph_upload(b200_ctx *ctx, int64_t n, int64_t m, const uint32_t *h_row_offsets,
                           const int32_t *h_col_indices, const float *h_col_values, b200_host_graph **out) {
    if (!ctx || !out || n < 1 || m < 0 || m >= (1ll << 32) || !h_row_offsets || (m && !h_col_indices)) return B200_ERR_INVALID;
    *out = nullptr;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    b200_host_graph *hg = new b200_host_graph;
    std::memset(hg, 0, sizeof(*hg));
    hg->n = n;
    hg->m = m;
    cudaStream_t st = ws_stream(&ctx->ws);
    int s = B200_OK;
    do {
        if ((s = cuda_status(cudaMalloc(&hg->d_row_offsets, sizeof(uint32_t) * (size_t)(n + 1))))) break;
        if ((s = cuda_status(cudaMalloc(&hg->d_col_indices, sizeof(int32_t) * (size_t)((m + 3) / 4 * 4 + 4))))) break;
        if (h_col_values && (s = cuda_status(cudaMalloc(&hg->d_col_values, sizeof(float) * (size_t)((m + 3) / 4 * 4 + 4))))) break;
        if ((s = cuda_status(cudaMalloc(&hg->d_labels, sizeof(int32_t) * (size_t)n)))) break;
        if ((s = cuda_status(cudaMalloc(&hg->d_dist, sizeof(float) * (size_t)n)))) break;
        if ((s = cuda_status(cudaMalloc(&hg->d_preds, sizeof(int32_t) * (size_t)n)))) break;
        if ((s = cuda_status(cudaMemcpyAsync(hg->d_row_offsets, h_row_offsets, sizeof(uint32_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st)))) break;
        if (m && (s = cuda_status(cudaMemcpyAsync(hg->d_col_indices, h_col_indices, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, st)))) break;
        if (m && h_col_values && (s = cuda_status(cudaMemcpyAsync(hg->d_col_values, h_col_values, sizeof(float) * (size_t)m, cudaMemcpyHostToDevice, st)))) break;
        if ((s = cuda_status(cudaMalloc(&hg->d_no_in_arc, sizeof(uint32_t) * (size_t)((n + 31) / 32 + 1))))) break;
        if ((s = cuda_status(launch_no_in_arc_bitmap(&ctx->ws, hg->d_row_offsets, n, hg->d_no_in_arc)))) break;
        if ((s = cuda_status(cudaMalloc(&hg->d_first_in_nbr, sizeof(int32_t) * (size_t)n)))) break;
        if ((s = cuda_status(launch_first_in_neighbor(&ctx->ws, hg->d_row_offsets, hg->d_col_indices, n, hg->d_first_in_nbr)))) break;
        s = cuda_status(cudaStreamSynchronize(st));
    } while (0);
    if (s != B200_OK) {
        b200_host_graph_free(ctx, hg);
        return s;
    }
    *out = hg;
    return B200_OK;
}

int b200_host_graph_free(b200_ctx *ctx, b200_host_graph *hg) {
    if (!hg) return B200_OK;
    if (ctx) cudaSetDevice(ctx->ws.device);
    cudaFree(hg->d_row_offsets);
    cudaFree(hg->d_col_indices);
    cudaFree(hg->d_col_values);
    cudaFree(hg->d_labels);
    cudaFree(hg->d_dist);
    cudaFree(hg->d_preds);
    cudaFree(hg->d_no_in_arc);
    cudaFree(hg->d_first_in_nbr);
    delete hg;
    return B200_OK;
}

int b200_host_graph_view(const b200_host_graph *hg, b200_graph *out) {
    if (!hg || !out) return B200_ERR_INVALID;
    out->n = hg->n;
    out->m = hg->m;
    out->row_offsets = hg->d_row_offsets;
    out->col_indices = hg->d_col_indices;
    out->col_values = hg->d_col_values;
    out->col_offsets = hg->d_row_offsets;   // symmetric graphs: CSC aliases CSR (graph.hxx:75-80)
    out->row_indices = hg->d_col_indices;
    out->row_values = hg->d_col_values;
    out->no_in_arc_bitmap = hg->d_no_in_arc;
    out->first_in_neighbor = hg->d_first_in_nbr;
    out->hot_ids = nullptr;
    out->hot_indices = nullptr;
    out->hot_count = 0;
    return B200_OK;
}

int b200_bfs_host(b200_ctx *ctx, b200_host_graph *hg, int32_t src, int mode, float alpha, float beta,
                  const int32_t *h_labels_init, int32_t *h_labels_out, b200_stats *stats) {
    if (!ctx || !hg || !h_labels_out) return B200_ERR_INVALID;
    b200_graph g;
    B200_TRY(b200_host_graph_view(hg, &g));
    cudaStream_t st = ws_stream(&ctx->ws);
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    // bfs_problem_t ctor uploads the host-initialised labels (bfs_problem.hxx:38-43: to_mem(labels));
    // the engine rebuilds them on the device anyway, the copy is kept so the timed region moves the
    // same bytes the reference's constructor does.
    if (h_labels_init)
        B200_CUDA(cudaMemcpyAsync(hg->d_labels, h_labels_init, sizeof(int32_t) * (size_t)hg->n, cudaMemcpyHostToDevice, st));
    B200_TRY(b200_bfs_run(ctx, &g, src, mode, alpha, beta, hg->d_labels, stats));
    // extract() (bfs_problem.hxx:48-50)
    B200_CUDA(cudaMemcpyAsync(h_labels_out, hg->d_labels, sizeof(int32_t) * (size_t)hg->n, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return B200_OK;
}

int b200_sssp_host(b200_ctx *ctx, b200_host_graph *hg, int32_t src, const float *h_dist_init, float *h_dist_out,
                   int32_t *h_preds_out, b200_stats *stats) {
    if (!ctx || !hg || !h_dist_out || !hg->d_col_values) return B200_ERR_INVALID;
    b200_graph g;
    B200_TRY(b200_host_graph_view(hg, &g));
    cudaStream_t st = ws_stream(&ctx->ws);
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    if (h_dist_init)
        B200_CUDA(cudaMemcpyAsync(hg->d_dist, h_dist_init, sizeof(float) * (size_t)hg->n, cudaMemcpyHostToDevice, st));
    B200_TRY(b200_sssp_run(ctx, &g, src, hg->d_dist, h_preds_out ? hg->d_preds : nullptr, stats));
    // extract() (sssp_problem.hxx:54-57)
    B200_CUDA(cudaMemcpyAsync(h_dist_out, hg->d_dist, sizeof(float) * (size_t)hg->n, cudaMemcpyDeviceToHost, st));
    if (h_preds_out)
        B200_CUDA(cudaMemcpyAsync(h_preds_out, hg->d_preds, sizeof(int32_t) * (size_t)hg->n, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return B200_OK;
}

}  // extern "C"
