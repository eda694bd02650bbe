// multi_gpu.cu -- per-rank steps of the 1D-partitioned (cyclic) multi-GPU BFS.
// The reference has no multi-GPU path (README.md:4); this is new.  One process
// drives one GPU; the exchanges between the steps below are NCCL collectives issued by the
// host layer (mini_b200/dist.py) on the same stream:
//   push level:  b200_mg_bfs_push   -> alltoall(counts) + alltoallv(vertex ids) -> b200_mg_bfs_absorb
//   pull level:  allgather(frontier bitmap slices) -> b200_mg_bfs_pull
// All buffers are caller-owned device memory described by b200_mg_bfs_state.
#include <climits>
#include "b200/operators.cuh"
#include "engine.cuh"

using namespace b200;

namespace {

__global__ void mg_init_kernel(int32_t *labels, uint32_t *known, int32_t *frontier, int src, Partition part) {
    const uint32_t b = part.bit((uint32_t)src);
    known[b >> 5] |= 1u << (b & 31);              // every rank knows the source is visited
    if (part.owner((uint32_t)src) == part.me) {
        labels[part.row((uint32_t)src)] = 0;
        frontier[0] = src;
    }
}

// local frontier list (global ids, all owned) -> this rank's bitmap slice
__global__ void mg_list_to_slice_kernel(const int32_t *list, uint32_t len, uint32_t *slice, Partition part) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) {
        const uint32_t r = part.row((uint32_t)list[i]);
        atomicOr(slice + (r >> 5), 1u << (r & 31));
    }
}

struct SlicePred {   // bit r of the slice set -> emit the global id of local row r
    const uint32_t *slice;
    Partition part;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = (int)part.global_id(part.me, idx);
        return (slice[idx >> 5] >> (idx & 31)) & 1u;
    }
};

int valid_state(const b200_mg_bfs_state *s) {
    if (!s || s->num_ranks < 1 || s->num_ranks > MAX_DEST || (s->num_ranks & (s->num_ranks - 1))) return 0;
    if (s->rank < 0 || s->rank >= s->num_ranks || s->n_local < 32 || s->n_local % 32) return 0;
    if (s->n_global != s->n_local * s->num_ranks) return 0;
    return s->labels && s->known && s->frontier_bitmap && s->next_slice && s->box_counts;
}

Partition part_of(const b200_mg_bfs_state *s) {
    uint32_t log_p = 0;
    while ((1 << log_p) < s->num_ranks) ++log_p;
    return Partition{log_p, (uint32_t)s->rank, (uint32_t)s->n_local};
}

}  // namespace

extern "C" {

int b200_partition_locate(int num_ranks, int64_t v, int32_t *owner, int64_t *row) {
    if (num_ranks < 1 || num_ranks > MAX_DEST || (num_ranks & (num_ranks - 1)) || v < 0 || v >= (1ll << 32)) return B200_ERR_INVALID;
    uint32_t log_p = 0;
    while ((1 << log_p) < num_ranks) ++log_p;
    const Partition part{log_p, 0u, 0u};
    if (owner) *owner = (int32_t)part.owner((uint32_t)v);
    if (row) *row = (int64_t)part.row((uint32_t)v);
    return B200_OK;
}

int b200_partition_global_id(int num_ranks, int32_t rank, int64_t row, int64_t *v) {
    if (num_ranks < 1 || num_ranks > MAX_DEST || (num_ranks & (num_ranks - 1)) || rank < 0 || rank >= num_ranks || !v) return B200_ERR_INVALID;
    uint32_t log_p = 0;
    while ((1 << log_p) < num_ranks) ++log_p;
    if (row < 0 || row >= (1ll << (32 - log_p))) return B200_ERR_INVALID;
    const Partition part{log_p, 0u, 0u};
    *v = (int64_t)part.global_id((uint32_t)rank, (uint32_t)row);
    return B200_OK;
}

int b200_mg_bfs_init(b200_ctx *ctx, const b200_mg_bfs_state *s, int32_t src, int32_t *d_frontier, int64_t *frontier_len) {
    if (!ctx || !valid_state(s) || !d_frontier || !frontier_len || src < 0 || src >= s->n_global) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    B200_TRY(b200_ctx_reserve(ctx, s->n_local));
    cudaStream_t st = ws_stream(&ctx->ws);
    const Partition part = part_of(s);
    B200_CUDA(cudaMemsetAsync(s->labels, 0xFF, sizeof(int32_t) * (size_t)s->n_local, st));
    B200_CUDA(cudaMemsetAsync(s->known, 0, sizeof(uint32_t) * (size_t)(s->n_global / 32), st));
    mg_init_kernel<<<1, 1, 0, st>>>(s->labels, s->known, d_frontier, src, part);
    ctx->ws.launches++;
    B200_CUDA(cudaGetLastError());
    *frontier_len = part.owner((uint32_t)src) == part.me ? 1 : 0;
    return B200_OK;
}

int b200_mg_bfs_push(b200_ctx *ctx, const b200_graph *g, const b200_mg_bfs_state *s, int level, const int32_t *d_frontier,
                     int64_t frontier_len, int32_t *d_next_frontier, int32_t *const *d_send_boxes, int64_t box_capacity,
                     int64_t *h_counts, int64_t *arcs, int64_t *next_degree) {
    if (!ctx || !g || !valid_state(s) || !d_next_frontier || !h_counts || frontier_len < 0 || (frontier_len && !d_frontier))
        return B200_ERR_INVALID;
    if (s->num_ranks > 1 && !d_send_boxes) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    b200_workspace *ws = &ctx->ws;
    cudaStream_t st = ws_stream(ws);
    const Partition part = part_of(s);
    B200_CUDA(reset_counters(ws));
    B200_CUDA(cudaMemsetAsync(s->box_counts, 0, sizeof(unsigned long long) * MAX_DEST, st));
    bool quad = false;
    if (frontier_len) {
        B200_TRY(b200_ctx_reserve(ctx, frontier_len));
        RoutedOut r;
        memset(&r, 0, sizeof r);
        r.num_dest = s->num_ranks;
        r.count = s->box_counts;
        for (int p = 0; p < s->num_ranks; ++p) {
            r.box[p] = p == s->rank ? d_next_frontier : d_send_boxes[p];
            r.capacity[p] = p == s->rank ? (unsigned long long)s->n_local : (unsigned long long)box_capacity;
        }
        quad = ctx->adv_impl != B200_ADVANCE_LBS && quad_aligned(g->col_indices);
        if (quad) {
            B200_CUDA(launch_quad_scan(ws, d_frontier, (uint32_t)frontier_len, g->row_offsets, part.log_p));
            const QuadArgs a = make_quad_args(ws, d_frontier, (uint32_t)frontier_len, g->row_offsets, g->col_indices, nullptr, part.log_p);
            BfsPushPartQ op{s->known, s->labels, level + 1, part};
            B200_CUDA((launch_quad_advance<OUT_ROUTED, true>(ws, a, op, nullptr, 0ull, &r)));
        } else {
            B200_CUDA(launch_frontier_scan(ws, d_frontier, (uint32_t)frontier_len, g->row_offsets, part.log_p));
            const LbsArgs a = make_lbs_args(ws, d_frontier, (uint32_t)frontier_len, g->row_offsets, g->col_indices, part.log_p);
            BfsPushPartOp op{s->known, s->labels, level + 1, part};
            B200_CUDA((launch_lbs_advance<OUT_ROUTED, true>(ws, a, op, nullptr, 0ull, &r)));
        }
    }
    unsigned long long h_box[MAX_DEST];
    B200_CUDA(cudaMemcpyAsync(h_box, s->box_counts, sizeof h_box, cudaMemcpyDeviceToHost, st));
    B200_CUDA(read_counters(ws));
    for (int p = 0; p < s->num_ranks; ++p) h_counts[p] = (int64_t)h_box[p];
    if (arcs) *arcs = (int64_t)ws->h_counters[quad ? B200_CNT_ARCS : B200_CNT_TOTAL];
    if (next_degree) *next_degree = (int64_t)ws->h_counters[B200_CNT_AUX];
    return ws->h_counters[B200_CNT_OVERFLOW] ? B200_ERR_OVERFLOW : B200_OK;
}

int b200_mg_bfs_absorb(b200_ctx *ctx, const b200_graph *g, const b200_mg_bfs_state *s, int level, const int32_t *d_inbox,
                       int64_t count, int32_t *d_next_frontier, int64_t *next_len, int64_t *next_degree) {
    if (!ctx || !g || !valid_state(s) || !d_next_frontier || !next_len || count < 0 || (count && !d_inbox)) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    b200_workspace *ws = &ctx->ws;
    cudaStream_t st = ws_stream(ws);
    const Partition part = part_of(s);
    B200_CUDA(reset_counters(ws));
    if (count) {
        bfs_absorb_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d_inbox, (uint32_t)count, s->known, s->labels, level + 1,
                                                                           part, d_next_frontier, (unsigned long long)s->n_local,
                                                                           s->box_counts + s->rank, ws->d_counters, g->row_offsets);
        ws->launches++;
        B200_CUDA(cudaGetLastError());
    }
    unsigned long long h = 0;
    B200_CUDA(cudaMemcpyAsync(&h, s->box_counts + s->rank, sizeof h, cudaMemcpyDeviceToHost, st));
    B200_CUDA(read_counters(ws));
    *next_len = (int64_t)h;
    if (next_degree) *next_degree = (int64_t)ws->h_counters[B200_CNT_AUX];
    return ws->h_counters[B200_CNT_OVERFLOW] ? B200_ERR_OVERFLOW : B200_OK;
}

int b200_mg_bfs_pull(b200_ctx *ctx, const b200_graph *g, const b200_mg_bfs_state *s, int level, int64_t *found,
                     int64_t *arcs_inspected, int64_t *found_degree) {
    if (!ctx || !g || !valid_state(s)) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    b200_workspace *ws = &ctx->ws;
    const Partition part = part_of(s);
    const uint32_t *off = g->col_offsets ? g->col_offsets : g->row_offsets;
    const int32_t *idx = g->row_indices ? g->row_indices : g->col_indices;
    B200_CUDA(reset_counters(ws));
    uint32_t *known_slice = s->known + (size_t)s->rank * (size_t)(s->n_local / 32);
    bfs_pull_kernel<256><<<ws->num_sms * 8, 256, 0, ws_stream(ws)>>>((uint32_t)s->n_local, off, idx, s->frontier_bitmap, s->next_slice,
                                                                      known_slice, s->labels, level + 1, ws->d_counters, part);
    ws->launches++;
    B200_CUDA(cudaGetLastError());
    B200_CUDA(read_counters(ws));
    if (found) *found = (int64_t)ws->h_counters[B200_CNT_OUT];
    if (arcs_inspected) *arcs_inspected = (int64_t)ws->h_counters[B200_CNT_ARCS];
    if (found_degree) *found_degree = (int64_t)ws->h_counters[B200_CNT_AUX];
    return B200_OK;
}

int b200_mg_bitmap_or(b200_ctx *ctx, uint32_t *d_dst, const uint32_t *d_src, int64_t words) {
    if (!ctx || !d_dst || !d_src || words < 0) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    if (words) {
        bitmap_or_kernel<<<ctx->ws.num_sms * 4, 256, 0, ws_stream(&ctx->ws)>>>(d_dst, d_src, (size_t)words);
        ctx->ws.launches++;
        B200_CUDA(cudaGetLastError());
    }
    return B200_OK;
}

int b200_mg_list_to_slice(b200_ctx *ctx, const b200_mg_bfs_state *s, const int32_t *d_list, int64_t len, uint32_t *d_slice) {
    if (!ctx || !valid_state(s) || !d_slice || len < 0 || (len && !d_list)) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    cudaStream_t st = ws_stream(&ctx->ws);
    B200_CUDA(cudaMemsetAsync(d_slice, 0, sizeof(uint32_t) * (size_t)(s->n_local / 32), st));
    if (len) {
        mg_list_to_slice_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(d_list, (uint32_t)len, d_slice, part_of(s));
        ctx->ws.launches++;
        B200_CUDA(cudaGetLastError());
    }
    return B200_OK;
}

int b200_mg_slice_to_list(b200_ctx *ctx, const b200_mg_bfs_state *s, const uint32_t *d_slice, int32_t *d_list, int64_t *len) {
    if (!ctx || !valid_state(s) || !d_slice || !d_list || !len) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    b200_workspace *ws = &ctx->ws;
    B200_TRY(b200_ctx_reserve(ctx, s->n_local));
    B200_CUDA(reset_counters(ws));
    B200_CUDA(launch_compact(ws, SlicePred{d_slice, part_of(s)}, (uint32_t)s->n_local, d_list, (unsigned long long)s->n_local,
                             ws->d_counters + B200_CNT_OUT, ws->d_counters + B200_CNT_OVERFLOW));
    B200_CUDA(read_counters(ws));
    *len = (int64_t)ws->h_counters[B200_CNT_OUT];
    return B200_OK;
}

}  // extern "C"
