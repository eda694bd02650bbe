// context.cu -- b200_ctx: stream, scratch arena, counters, L2 persistence.
// Replaces the role of mgpu::standard_context_t (context.hxx:103-219) and all the
// per-operator-call mem_t scratch of the reference.
#include <cstdio>
#include <cstring>
#include <new>
#include "engine.cuh"
#include "b200/device_utils.cuh"
#include "b200/loop_dyn.cuh"

namespace b200 {
thread_local int g_last_cuda_error = 0;

static void free_traversal_scratch(b200_ctx *ctx) {
    level_loop_invalidate(ctx);
    ctx->scratch_gen++;
    for (int i = 0; i < 2; ++i) {
        if (ctx->frontier[i]) cudaFree(ctx->frontier[i]);
        if (ctx->bm_frontier[i]) cudaFree(ctx->bm_frontier[i]);
        ctx->frontier[i] = nullptr;
        ctx->bm_frontier[i] = nullptr;
    }
    if (ctx->bm_visited) cudaFree(ctx->bm_visited);
    if (ctx->stamp) cudaFree(ctx->stamp);
    ctx->bm_visited = nullptr;
    ctx->stamp = nullptr;
    ctx->scratch_n = 0;
}

static __global__ void isolated_bitmap_kernel(const uint32_t *__restrict__ offsets, uint32_t n, uint32_t num_words, uint32_t *bm) {
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t word = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; word < num_words; word += warps_total) {
        const uint32_t v = (word << 5) + lane;
        const bool iso = v < n && offsets[v + 1] == offsets[v];
        const unsigned mask = __ballot_sync(0xffffffffu, iso);
        if (lane == 0) bm[word] = mask;
    }
}

// visited |= "vertex has no in-arc" (iso: the caller's prebuilt bitmap, else derived from the offsets)
static __global__ void or_no_in_arc_kernel(const uint32_t *__restrict__ offsets, uint32_t n, const uint32_t *__restrict__ iso,
                                           uint32_t *visited) {
    or_no_in_arc_words(blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, offsets, n, iso, visited);
}

static __global__ void first_in_neighbor_kernel(const uint32_t *__restrict__ offsets, const int32_t *__restrict__ indices,
                                                unsigned long long n, int32_t *__restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const uint32_t b = offsets[v], e = offsets[v + 1];
        out[v] = e > b ? indices[b] : -1;
    }
}

cudaError_t launch_first_in_neighbor(b200_workspace *ws, const uint32_t *pull_offsets, const int32_t *pull_indices, int64_t n,
                                     int32_t *d_out) {
    first_in_neighbor_kernel<<<ws->num_sms * 8, 256, 0, (cudaStream_t)ws->stream>>>(pull_offsets, pull_indices,
                                                                                    (unsigned long long)n, d_out);
    ws->launches++;
    return cudaGetLastError();
}

cudaError_t preload_no_in_arc_kernel() {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, isolated_bitmap_kernel);
}

cudaError_t launch_no_in_arc_bitmap(b200_workspace *ws, const uint32_t *pull_offsets, int64_t n, uint32_t *d_bitmap) {
    const int64_t words = (n + 31) / 32;
    isolated_bitmap_kernel<<<ws->num_sms * 8, 256, 0, (cudaStream_t)ws->stream>>>(pull_offsets, (uint32_t)n, (uint32_t)words, d_bitmap);
    ws->launches++;
    return cudaGetLastError();
}

cudaError_t launch_or_no_in_arc(b200_workspace *ws, const uint32_t *pull_offsets, int64_t n, const uint32_t *iso, uint32_t *d_visited) {
    or_no_in_arc_kernel<<<ws->num_sms * 8, 256, 0, (cudaStream_t)ws->stream>>>(pull_offsets, (uint32_t)n, iso, d_visited);
    ws->launches++;
    return cudaGetLastError();
}

int ensure_traversal_scratch(b200_ctx *ctx, int64_t n) {
    if (n <= ctx->scratch_n) return B200_OK;
    B200_CUDA(cudaStreamSynchronize((cudaStream_t)ctx->ws.stream));
    free_traversal_scratch(ctx);
    const size_t words = (size_t)((n + 31) / 32) + 1;
    for (int i = 0; i < 2; ++i) {
        B200_CUDA(cudaMalloc(&ctx->frontier[i], sizeof(int32_t) * (size_t)n));
        B200_CUDA(cudaMalloc(&ctx->bm_frontier[i], sizeof(uint32_t) * words));
    }
    B200_CUDA(cudaMalloc(&ctx->bm_visited, sizeof(uint32_t) * words));
    B200_CUDA(cudaMalloc(&ctx->stamp, sizeof(int32_t) * (size_t)n));
    ctx->scratch_n = n;
    return b200_ctx_reserve(ctx, n);
}
}  // namespace b200

using namespace b200;

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }

const char *b200_status_string(int status) {
    switch (status) {
        case B200_OK: return "ok";
        case B200_ERR_CUDA: return "CUDA runtime error (see b200_last_cuda_error)";
        case B200_ERR_INVALID: return "invalid argument";
        case B200_ERR_OVERFLOW: return "output frontier capacity exceeded";
        case B200_ERR_NOMEM: return "out of device memory";
        case B200_ERR_UNSUPPORTED: return "unsupported";
        case B200_ERR_TIMEOUT: return "a peer GPU did not reach a barrier in time";
        case B200_ERR_IO: return "file could not be opened, read or written";
        case B200_ERR_FORMAT: return "malformed file";
        default: return "unknown status";
    }
}

int b200_last_cuda_error(void) { return g_last_cuda_error; }

int b200_device_count(int *count) {
    if (!count) return B200_ERR_INVALID;
    *count = 0;
    B200_CUDA(cudaGetDeviceCount(count));
    return B200_OK;
}

b200_workspace *b200_ctx_workspace(b200_ctx *ctx) { return ctx ? &ctx->ws : nullptr; }

int b200_ctx_create(b200_ctx **out, int device, void *stream) {
    if (!out) return B200_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    B200_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(device));
    b200_ctx *ctx = new (std::nothrow) b200_ctx;
    if (!ctx) return B200_ERR_NOMEM;
    std::memset(ctx, 0, sizeof(*ctx));
    ctx->ws.device = device;
    B200_CUDA(cudaDeviceGetAttribute(&ctx->ws.num_sms, cudaDevAttrMultiProcessorCount, device));
    if (stream) {
        ctx->ws.stream = stream;
        ctx->own_stream = false;
    } else {
        cudaStream_t s;
        B200_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        ctx->ws.stream = (void *)s;
        ctx->own_stream = true;
    }
    B200_CUDA(cudaMalloc(&ctx->ws.d_counters, sizeof(unsigned long long) * B200_NUM_COUNTERS));
    B200_CUDA(cudaMemset(ctx->ws.d_counters, 0, sizeof(unsigned long long) * B200_NUM_COUNTERS));
    B200_CUDA(cudaMallocHost(&ctx->ws.h_counters, sizeof(unsigned long long) * B200_NUM_COUNTERS));
    B200_CUDA(cudaMalloc(&ctx->ws.d_tile_counter, 64));
    B200_CUDA(cudaMalloc(&ctx->d_nf, sizeof(b200::NearFar)));
    B200_CUDA(cudaMemset(ctx->d_nf, 0, sizeof(b200::NearFar)));
    B200_CUDA(cudaMalloc(&ctx->d_nf_sum, sizeof(double)));
    B200_CUDA(cudaMemset(ctx->ws.d_tile_counter, 0, 64));
    B200_CUDA(cudaEventCreate(&ctx->ev_run[0]));
    B200_CUDA(cudaEventCreate(&ctx->ev_run[1]));
    *out = ctx;
    return b200_ctx_reserve(ctx, 1 << 20);
}

int b200_ctx_reserve(b200_ctx *ctx, int64_t max_items) {
    if (!ctx || max_items < 0) return B200_ERR_INVALID;
    if (max_items >= (1ll << 32)) return B200_ERR_UNSUPPORTED;
    b200_workspace &ws = ctx->ws;
    B200_CUDA(cudaSetDevice(ws.device));
    // smallest tile any scan-type kernel uses is 1024 items
    const int64_t tiles = (max_items + 1023) / 1024 + 1;
    if (tiles > ws.status_tiles || max_items > ws.scanned_capacity) {
        level_loop_invalidate(ctx);
        ctx->scratch_gen++;
    }
    if (tiles > ws.status_tiles) {
        B200_CUDA(cudaStreamSynchronize((cudaStream_t)ws.stream));
        if (ws.d_status) cudaFree(ws.d_status);
        ws.d_status = nullptr;
        ws.status_tiles = 0;
        B200_CUDA(cudaMalloc(&ws.d_status, sizeof(unsigned long long) * (size_t)tiles));
        B200_CUDA(cudaMemset(ws.d_status, 0, sizeof(unsigned long long) * (size_t)tiles));
        ws.status_tiles = tiles;
        ws.epoch = 0;
    }
    if (max_items > ws.scanned_capacity) {
        B200_CUDA(cudaStreamSynchronize((cudaStream_t)ws.stream));
        if (ws.d_scanned) cudaFree(ws.d_scanned);
        if (ws.d_rows) cudaFree(ws.d_rows);
        if (ws.d_scanned2) cudaFree(ws.d_scanned2);
        if (ws.d_rows2) cudaFree(ws.d_rows2);
        ws.d_scanned = ws.d_scanned2 = nullptr;
        ws.d_rows = ws.d_rows2 = nullptr;
        ws.scanned_capacity = 0;
        B200_CUDA(cudaMalloc(&ws.d_scanned, sizeof(uint32_t) * (size_t)(max_items + 1)));
        B200_CUDA(cudaMalloc(&ws.d_rows, 2 * sizeof(uint32_t) * (size_t)(max_items + 1)));
        B200_CUDA(cudaMalloc(&ws.d_scanned2, sizeof(uint32_t) * (size_t)(max_items + 1)));
        B200_CUDA(cudaMalloc(&ws.d_rows2, 2 * sizeof(uint32_t) * (size_t)(max_items + 1)));
        ws.scanned_capacity = max_items;
    }
    return B200_OK;
}

int b200_ctx_sync(b200_ctx *ctx) {
    if (!ctx) return B200_ERR_INVALID;
    B200_CUDA(cudaStreamSynchronize((cudaStream_t)ctx->ws.stream));
    return B200_OK;
}

int b200_ctx_num_sms(b200_ctx *ctx, int *num_sms) {
    if (!ctx || !num_sms) return B200_ERR_INVALID;
    *num_sms = ctx->ws.num_sms;
    return B200_OK;
}

int b200_ctx_set_advance_impl(b200_ctx *ctx, int impl) {
    if (!ctx || (impl != B200_ADVANCE_QUAD && impl != B200_ADVANCE_LBS && impl != B200_ADVANCE_QUAD_RESCAN)) return B200_ERR_INVALID;
    ctx->adv_impl = impl;
    return B200_OK;
}

int b200_ctx_set_level_loop(b200_ctx *ctx, int impl) {
    if (!ctx || (impl != B200_LOOP_GRAPH && impl != B200_LOOP_HOST)) return B200_ERR_INVALID;
    ctx->loop_impl = impl;
    return B200_OK;
}

int b200_ctx_forget_graph(b200_ctx *ctx) {
    if (!ctx) return B200_ERR_INVALID;
    level_loop_invalidate(ctx);
    ctx->nf_key_w = nullptr;
    return B200_OK;
}

int b200_ctx_set_sssp_delta(b200_ctx *ctx, float delta) {
    if (!ctx || !(delta >= 0.f)) return B200_ERR_INVALID;
    ctx->sssp_delta = delta;
    return B200_OK;
}

int b200_graph_first_in_neighbor(b200_ctx *ctx, const b200_graph *g, int32_t *d_out) {
    if (!ctx || !g || !d_out || g->n < 1 || g->n > (1ll << 31)) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    B200_CUDA(launch_first_in_neighbor(&ctx->ws, g->col_offsets ? g->col_offsets : g->row_offsets,
                                       g->row_indices ? g->row_indices : g->col_indices, g->n, d_out));
    return B200_OK;
}

int b200_graph_no_in_arc_bitmap(b200_ctx *ctx, const b200_graph *g, uint32_t *d_bitmap) {
    if (!ctx || !g || !d_bitmap || g->n < 1 || g->n > (1ll << 31)) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    B200_CUDA(launch_no_in_arc_bitmap(&ctx->ws, g->col_offsets ? g->col_offsets : g->row_offsets, g->n, d_bitmap));
    return B200_OK;
}

int b200_ctx_l2_pin(b200_ctx *ctx, const void *d_ptr, int64_t bytes) {
    if (!ctx || bytes < 0) return B200_ERR_INVALID;
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    if (bytes == 0 || !d_ptr) {
        attr.accessPolicyWindow.num_bytes = 0;
        B200_CUDA(cudaStreamSetAttribute((cudaStream_t)ctx->ws.stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        if (ctx->l2_window_set) (void)cudaCtxResetPersistingL2Cache();
        ctx->l2_window_set = false;
        return B200_OK;
    }
    int max_window = 0, max_persist = 0;
    B200_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->ws.device));
    B200_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->ws.device));
    if (max_window <= 0 || max_persist <= 0) return B200_ERR_UNSUPPORTED;
    size_t win = (size_t)bytes < (size_t)max_window ? (size_t)bytes : (size_t)max_window;
    size_t carve = win < (size_t)max_persist ? win : (size_t)max_persist;
    B200_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(d_ptr);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = (float)((double)carve / (double)win);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    B200_CUDA(cudaStreamSetAttribute((cudaStream_t)ctx->ws.stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    ctx->l2_window_set = true;
    return B200_OK;
}

int b200_ctx_destroy(b200_ctx *ctx) {
    if (!ctx) return B200_OK;
    cudaSetDevice(ctx->ws.device);
    cudaStreamSynchronize((cudaStream_t)ctx->ws.stream);
    if (ctx->l2_window_set) b200_ctx_l2_pin(ctx, nullptr, 0);
    level_loop_destroy(ctx);
    free_traversal_scratch(ctx);
    if (ctx->ev_level) {
        for (int i = 0; i < ctx->ev_level_count; ++i) cudaEventDestroy(ctx->ev_level[i]);
        delete[] ctx->ev_level;
    }
    cudaEventDestroy(ctx->ev_run[0]);
    cudaEventDestroy(ctx->ev_run[1]);
    if (ctx->ws.d_status) cudaFree(ctx->ws.d_status);
    if (ctx->ws.d_scanned) cudaFree(ctx->ws.d_scanned);
    if (ctx->ws.d_rows) cudaFree(ctx->ws.d_rows);
    if (ctx->ws.d_scanned2) cudaFree(ctx->ws.d_scanned2);
    if (ctx->ws.d_rows2) cudaFree(ctx->ws.d_rows2);
    if (ctx->ws.d_counters) cudaFree(ctx->ws.d_counters);
    if (ctx->ws.h_counters) cudaFreeHost(ctx->ws.h_counters);
    if (ctx->ws.d_tile_counter) cudaFree(ctx->ws.d_tile_counter);
    if (ctx->hot_vals) cudaFree(ctx->hot_vals);
    if (ctx->d_nf) cudaFree(ctx->d_nf);
    if (ctx->d_nf_sum) cudaFree(ctx->d_nf_sum);
    if (ctx->own_stream) cudaStreamDestroy((cudaStream_t)ctx->ws.stream);
    delete ctx;
    return B200_OK;
}

}  // extern "C"
