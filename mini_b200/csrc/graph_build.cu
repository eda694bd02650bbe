// graph_build.cu -- device-side synthetic RMAT generator and CSR builder.
// Produces what graph.hxx:139-172 builds from a sorted tuple list (arcs ordered by
// (src,dst), duplicates and self loops kept, empty rows get the next offset) and
// what load_graph(_, true, _) does for symmetrisation (graph.hxx:130-137), but on
// the GPU and with offsets that can hold 2^31 arcs.  Not on the timed hot path:
// uses cub::DeviceRadixSort (CUDA toolkit library) for the (src,dst) sort.
#include <cub/device/device_radix_sort.cuh>
#include "b200/partition.cuh"
#include "engine.cuh"
#include "rmat.cuh"

using namespace b200;

namespace {

__global__ void rmat_keys_kernel(uint64_t key, int scale, unsigned long long npairs, unsigned long long *keys) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < npairs; e += stride) {
        uint32_t u, v;
        rmat_pair(key, e, scale, u, v);
        keys[2 * e] = ((unsigned long long)u << 32) | v;
        keys[2 * e + 1] = ((unsigned long long)v << 32) | u;
    }
}

__global__ void rmat_pairs_kernel(uint64_t key, int scale, unsigned long long npairs, int32_t *src, int32_t *dst) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < npairs; e += stride) {
        uint32_t u, v;
        rmat_pair(key, e, scale, u, v);
        src[e] = (int32_t)u;
        dst[e] = (int32_t)v;
    }
}

__global__ void pairs_to_keys_kernel(const int32_t *src, const int32_t *dst, unsigned long long npairs, int symmetrize,
                                     unsigned long long *keys) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < npairs; e += stride) {
        const uint32_t u = (uint32_t)src[e], v = (uint32_t)dst[e];
        if (symmetrize) {
            keys[2 * e] = ((unsigned long long)u << 32) | v;
            keys[2 * e + 1] = ((unsigned long long)v << 32) | u;
        } else {
            keys[e] = ((unsigned long long)u << 32) | v;
        }
    }
}

// sorted (src<<32|dst) keys -> col_indices, row_offsets (every row r gets the index of
// its first arc; empty rows inherit the next row's), optional pair weights.
__global__ void keys_to_csr_kernel(const unsigned long long *keys, unsigned long long m, unsigned long long n,
                                   uint32_t *row_offsets, int32_t *col_indices, float *weights, uint64_t wkey) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const unsigned long long k = keys[i];
        const uint32_t s = (uint32_t)(k >> 32), d = (uint32_t)k;
        col_indices[i] = (int32_t)d;
        if (weights) weights[i] = pair_weight(wkey, s, d);
        const long long prev = i ? (long long)(keys[i - 1] >> 32) : -1ll;
        for (long long r = prev + 1; r <= (long long)s; ++r) row_offsets[r] = (uint32_t)i;
        if (i == m - 1)
            for (unsigned long long r = (unsigned long long)s + 1; r <= n; ++r) row_offsets[r] = (uint32_t)m;
    }
}

// Swizzled-cyclic 1D partition (b200::Partition, advance.cuh): keep the arcs whose tail is owned by `rank`;
// key = (local row << 32) | global column.  count_only => just count them.
__global__ void rmat_part_keys_kernel(uint64_t key, int scale, unsigned long long npairs, uint32_t pmask, uint32_t log_p,
                                      uint32_t rank, unsigned long long *keys, unsigned long long *counter, int count_only) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounds = (npairs + stride - 1) / stride;
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long local = 0;
    const b200::Partition part{log_p, rank, 0u};
    for (unsigned long long r = 0; r < rounds; ++r, e += stride) {
        uint32_t u = 0, v = 0;
        bool fu = false, fv = false;
        if (e < npairs) {
            rmat_pair(key, e, scale, u, v);
            fu = part.owner(u) == rank;   // arc u -> v lives here
            fv = part.owner(v) == rank;   // arc v -> u lives here
        }
        if (count_only) {
            local += (fu ? 1 : 0) + (fv ? 1 : 0);
            continue;
        }
        const unsigned mu = __ballot_sync(0xffffffffu, fu), mv = __ballot_sync(0xffffffffu, fv);
        const unsigned total = __popc(mu) + __popc(mv);
        if (total) {
            unsigned long long base = 0;
            const unsigned lane = threadIdx.x & 31;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned lt = (1u << lane) - 1u;
            if (fu) keys[base + __popc(mu & lt)] = ((unsigned long long)(u >> log_p) << 32) | v;
            if (fv) keys[base + __popc(mu) + __popc(mv & lt)] = ((unsigned long long)(v >> log_p) << 32) | u;
        }
    }
    if (count_only) {
        for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, local);
    }
}

__global__ void fill_u32_kernel(uint32_t *p, unsigned long long count, uint32_t v) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) p[i] = v;
}

int sort_and_emit(b200_ctx *ctx, unsigned long long *keys_a, unsigned long long *keys_b, long long m, long long n,
                  int key_bits, uint32_t *d_row_offsets, int32_t *d_col_indices, float *d_weights, uint64_t wseed) {
    cudaStream_t st = (cudaStream_t)ctx->ws.stream;
    if (m == 0) {
        fill_u32_kernel<<<256, 256, 0, st>>>(d_row_offsets, (unsigned long long)n + 1, 0u);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaStreamSynchronize(st));
        return B200_OK;
    }
    cub::DoubleBuffer<unsigned long long> buf(keys_a, keys_b);
    size_t temp_bytes = 0;
    B200_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, buf, (long long)m, 0, key_bits, st));
    void *temp = nullptr;
    B200_CUDA(cudaMalloc(&temp, temp_bytes ? temp_bytes : 16));
    cudaError_t e = cub::DeviceRadixSort::SortKeys(temp, temp_bytes, buf, (long long)m, 0, key_bits, st);
    if (e == cudaSuccess) {
        keys_to_csr_kernel<<<ctx->ws.num_sms * 8, 256, 0, st>>>(buf.Current(), (unsigned long long)m, (unsigned long long)n,
                                                               d_row_offsets, d_col_indices, d_weights, weight_key(wseed));
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(temp);
    B200_CUDA(e);
    B200_CUDA(e2);
    return B200_OK;
}

// ---- b200_graph_hot_columns: the most frequent columns of an index array, and the remapped copy --------------
__global__ void hot_count_kernel(const int32_t *__restrict__ idx, unsigned long long m, uint32_t *counts) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += stride) atomicAdd(counts + idx[e], 1u);
}
__global__ void hot_keys_kernel(const uint32_t *__restrict__ counts, unsigned long long n, unsigned long long *keys) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;   // descending sort: more occurrences first, then smaller id
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        keys[v] = ((unsigned long long)counts[v] << 32) | (0xFFFFFFFFull - v);
}
__global__ void hot_slots_kernel(const unsigned long long *__restrict__ sorted, uint32_t hot_count, int32_t *hot_ids, int32_t *slot_of) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < hot_count) {
        const int32_t v = (int32_t)(0xFFFFFFFFull - (sorted[k] & 0xFFFFFFFFull));
        hot_ids[k] = v;
        slot_of[v] = (int32_t)k;
    }
}
__global__ void hot_remap_kernel(const int32_t *__restrict__ idx, unsigned long long m, const int32_t *__restrict__ slot_of, int32_t *out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += stride) {
        const int32_t v = idx[e];
        const int32_t k = slot_of[v];
        out[e] = k >= 0 ? ~k : v;
    }
}

}  // namespace

extern "C" {

int b200_graph_hot_columns(b200_ctx *ctx, const b200_graph *g, int64_t hot_count, int32_t *d_hot_ids, int32_t *d_hot_indices) {
    if (!ctx || !g || !d_hot_ids || !d_hot_indices || hot_count < 1 || hot_count > B200_HOT_MAX || hot_count > g->n || g->n < 1 ||
        g->n > (1ll << 31))
        return B200_ERR_INVALID;
    const int32_t *idx = g->row_indices ? g->row_indices : g->col_indices;   // the pull arrays (CSC aliases CSR: graph.hxx:75-80)
    if (!idx && g->m) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    cudaStream_t st = (cudaStream_t)ctx->ws.stream;
    const unsigned long long n = (unsigned long long)g->n, m = (unsigned long long)g->m;
    uint32_t *counts = nullptr;
    int32_t *slot_of = nullptr;
    unsigned long long *ka = nullptr, *kb = nullptr;
    void *temp = nullptr;
    int s = B200_OK;
    do {
        if ((s = cuda_status(cudaMalloc(&counts, sizeof(uint32_t) * n)))) break;
        if ((s = cuda_status(cudaMalloc(&slot_of, sizeof(int32_t) * n)))) break;
        if ((s = cuda_status(cudaMalloc(&ka, sizeof(unsigned long long) * n)))) break;
        if ((s = cuda_status(cudaMalloc(&kb, sizeof(unsigned long long) * n)))) break;
        if ((s = cuda_status(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * n, st)))) break;
        if ((s = cuda_status(cudaMemsetAsync(slot_of, 0xFF, sizeof(int32_t) * n, st)))) break;
        if (m) hot_count_kernel<<<ctx->ws.num_sms * 8, 256, 0, st>>>(idx, m, counts);
        hot_keys_kernel<<<ctx->ws.num_sms * 8, 256, 0, st>>>(counts, n, ka);
        if ((s = cuda_status(cudaGetLastError()))) break;
        cub::DoubleBuffer<unsigned long long> buf(ka, kb);
        size_t temp_bytes = 0;
        if ((s = cuda_status(cub::DeviceRadixSort::SortKeysDescending(nullptr, temp_bytes, buf, (long long)n, 0, 64, st)))) break;
        if ((s = cuda_status(cudaMalloc(&temp, temp_bytes ? temp_bytes : 16)))) break;
        if ((s = cuda_status(cub::DeviceRadixSort::SortKeysDescending(temp, temp_bytes, buf, (long long)n, 0, 64, st)))) break;
        hot_slots_kernel<<<(unsigned)((hot_count + 255) / 256), 256, 0, st>>>(buf.Current(), (uint32_t)hot_count, d_hot_ids, slot_of);
        if (m) hot_remap_kernel<<<ctx->ws.num_sms * 8, 256, 0, st>>>(idx, m, slot_of, d_hot_indices);
        if ((s = cuda_status(cudaGetLastError()))) break;
        s = cuda_status(cudaStreamSynchronize(st));
    } while (0);
    cudaFree(counts);
    cudaFree(slot_of);
    cudaFree(ka);
    cudaFree(kb);
    cudaFree(temp);
    return s;
}

int b200_rmat_pairs(b200_ctx *ctx, int scale, int edge_factor, uint64_t seed, int32_t *d_src, int32_t *d_dst) {
    if (!ctx || scale < 1 || scale > 31 || edge_factor < 1 || !d_src || !d_dst) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    const unsigned long long npairs = (unsigned long long)edge_factor << scale;
    rmat_pairs_kernel<<<ctx->ws.num_sms * 8, 256, 0, (cudaStream_t)ctx->ws.stream>>>(rmat_key(seed), scale, npairs, d_src, d_dst);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize((cudaStream_t)ctx->ws.stream));
    return B200_OK;
}

int b200_rmat_build_csr(b200_ctx *ctx, int scale, int edge_factor, uint64_t seed, uint32_t *d_row_offsets,
                        int32_t *d_col_indices, float *d_weights, uint64_t weight_seed) {
    if (!ctx || scale < 1 || scale > 31 || edge_factor < 1 || !d_row_offsets || !d_col_indices) return B200_ERR_INVALID;
    const unsigned long long npairs = (unsigned long long)edge_factor << scale;
    const unsigned long long m = 2 * npairs;
    if (m >= (1ull << 32)) return B200_ERR_UNSUPPORTED;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    unsigned long long *ka = nullptr, *kb = nullptr;
    B200_CUDA(cudaMalloc(&ka, sizeof(unsigned long long) * m));
    cudaError_t e = cudaMalloc(&kb, sizeof(unsigned long long) * m);
    if (e != cudaSuccess) { cudaFree(ka); return cuda_status(e); }
    rmat_keys_kernel<<<ctx->ws.num_sms * 8, 256, 0, (cudaStream_t)ctx->ws.stream>>>(rmat_key(seed), scale, npairs, ka);
    int s = cuda_status(cudaGetLastError());
    if (s == B200_OK)
        s = sort_and_emit(ctx, ka, kb, (long long)m, 1ll << scale, 32 + scale, d_row_offsets, d_col_indices, d_weights, weight_seed);
    cudaFree(ka);
    cudaFree(kb);
    return s;
}

static int part_args_ok(int scale, int edge_factor, int rank, int num_ranks) {
    if (scale < 1 || scale > 31 || edge_factor < 1) return 0;
    if (num_ranks < 1 || num_ranks > 8 || (num_ranks & (num_ranks - 1))) return 0;
    if (rank < 0 || rank >= num_ranks) return 0;
    return ((1ll << scale) / num_ranks) % 32 == 0;
}

int b200_rmat_part_count(b200_ctx *ctx, int scale, int edge_factor, uint64_t seed, int rank, int num_ranks, int64_t *m_local) {
    if (!ctx || !m_local || !part_args_ok(scale, edge_factor, rank, num_ranks)) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    cudaStream_t st = (cudaStream_t)ctx->ws.stream;
    uint32_t log_p = 0;
    while ((1 << log_p) < num_ranks) ++log_p;
    unsigned long long *cnt = ctx->ws.d_counters + B200_CNT_AUX2;
    B200_CUDA(cudaMemsetAsync(cnt, 0, sizeof(*cnt), st));
    rmat_part_keys_kernel<<<ctx->ws.num_sms * 8, 256, 0, st>>>(rmat_key(seed), scale, (unsigned long long)edge_factor << scale,
                                                               (uint32_t)num_ranks - 1, log_p, (uint32_t)rank, nullptr, cnt, 1);
    B200_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    B200_CUDA(cudaMemcpyAsync(&h, cnt, sizeof h, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    *m_local = (int64_t)h;
    return B200_OK;
}

int b200_rmat_build_csr_part(b200_ctx *ctx, int scale, int edge_factor, uint64_t seed, int rank, int num_ranks,
                             int64_t m_local, uint32_t *d_row_offsets, int32_t *d_col_indices) {
    if (!ctx || !d_row_offsets || !d_col_indices || m_local < 0 || m_local >= (1ll << 32) ||
        !part_args_ok(scale, edge_factor, rank, num_ranks))
        return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    cudaStream_t st = (cudaStream_t)ctx->ws.stream;
    uint32_t log_p = 0;
    while ((1 << log_p) < num_ranks) ++log_p;
    const unsigned long long m = (unsigned long long)m_local;
    unsigned long long *ka = nullptr, *kb = nullptr;
    B200_CUDA(cudaMalloc(&ka, sizeof(unsigned long long) * (m ? m : 1)));
    cudaError_t e = cudaMalloc(&kb, sizeof(unsigned long long) * (m ? m : 1));
    if (e != cudaSuccess) { cudaFree(ka); return cuda_status(e); }
    unsigned long long *cnt = ctx->ws.d_counters + B200_CNT_AUX2;
    int s = cuda_status(cudaMemsetAsync(cnt, 0, sizeof(*cnt), st));
    if (s == B200_OK) {
        rmat_part_keys_kernel<<<ctx->ws.num_sms * 8, 256, 0, st>>>(rmat_key(seed), scale, (unsigned long long)edge_factor << scale,
                                                                   (uint32_t)num_ranks - 1, log_p, (uint32_t)rank, ka, cnt, 0);
        s = cuda_status(cudaGetLastError());
    }
    unsigned long long h = 0;
    if (s == B200_OK) s = cuda_status(cudaMemcpyAsync(&h, cnt, sizeof h, cudaMemcpyDeviceToHost, st));
    if (s == B200_OK) s = cuda_status(cudaStreamSynchronize(st));
    if (s == B200_OK && h != m) s = B200_ERR_INVALID;   // m_local must come from b200_rmat_part_count
    if (s == B200_OK)
        s = sort_and_emit(ctx, ka, kb, (long long)m, (1ll << scale) / num_ranks, 32 + scale, d_row_offsets, d_col_indices, nullptr, 0);
    cudaFree(ka);
    cudaFree(kb);
    return s;
}

int b200_build_csr_from_pairs(b200_ctx *ctx, int64_t n, int64_t npairs, const int32_t *d_src, const int32_t *d_dst,
                              int symmetrize, uint32_t *d_row_offsets, int32_t *d_col_indices, float *d_weights,
                              uint64_t weight_seed) {
    if (!ctx || n < 1 || n > (1ll << 31) || npairs < 0 || !d_row_offsets || (npairs && (!d_src || !d_dst || !d_col_indices)))
        return B200_ERR_INVALID;
    const unsigned long long m = (unsigned long long)npairs * (symmetrize ? 2 : 1);
    if (m >= (1ull << 32)) return B200_ERR_UNSUPPORTED;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    unsigned long long *ka = nullptr, *kb = nullptr;
    B200_CUDA(cudaMalloc(&ka, sizeof(unsigned long long) * (m ? m : 1)));
    cudaError_t e = cudaMalloc(&kb, sizeof(unsigned long long) * (m ? m : 1));
    if (e != cudaSuccess) { cudaFree(ka); return cuda_status(e); }
    int s = B200_OK;
    if (npairs) {
        pairs_to_keys_kernel<<<ctx->ws.num_sms * 8, 256, 0, (cudaStream_t)ctx->ws.stream>>>(d_src, d_dst, (unsigned long long)npairs, symmetrize, ka);
        s = cuda_status(cudaGetLastError());
    }
    int bits = 1;
    while ((1ll << bits) < n) ++bits;
    if (s == B200_OK) s = sort_and_emit(ctx, ka, kb, (long long)m, n, 32 + bits, d_row_offsets, d_col_indices, d_weights, weight_seed);
    cudaFree(ka);
    cudaFree(kb);
    return s;
}

}  // extern "C"
