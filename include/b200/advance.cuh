// advance.cuh -- the advance kernels (push over LBS, pull over bitmaps).
//
// Push replaces the transform_lbs user lambda `neighbors_expand`
// (advance.hxx:53-61) AND the `!= -1` filter_kernel that always follows it
// (bfs_enactor.hxx:60-65, filter.hxx:11-31): accepted neighbours are appended
// to the output frontier through per-warp shared-memory stages (warp ballot ->
// stage -> one global atomic per ~256 accepted vertices -> coalesced copy), so the
// m_F-sized intermediate with -1 holes never exists.
// OUT_RAW keeps the reference layout out[idx] = nbr | -1.
//
// An Op supplies the per-arc functor in two halves so the kernel can batch the
// loads of all VT arcs of a thread before any atomic is issued:
//   bool probe (src, dst, edge_id)                       side-effect free, may over-approximate
//   bool commit(src, dst, edge_id, rank, out_idx)        atomics / writes; true => emit dst
#pragma once
#include "device_utils.cuh"
#include "lbs.cuh"
#include "workspace.h"

namespace b200 {

enum { OUT_NONE = 0, OUT_COMPACT = 1, OUT_RAW = 2 };

// DEG_SUM: also accumulate sum(deg(u)) over the emitted vertices into
// counters[B200_CNT_AUX] (one atomic per CTA) so the next level's m_F -- the input
// of the push/pull decision -- is known without another pass.
//
// Output staging is per WARP (WSTAGE ints of shared memory each): a warp appends its
// accepted neighbours with a ballot + popc, and when its stage is nearly full claims
// a range of the output frontier with ONE global atomic and copies the stage out with
// coalesced stores.  No CTA barrier is involved, so together with the barrier-free
// tile walk of lbs.cuh the warps of a CTA never wait for each other inside a window.
template <class Op, int OUT_MODE, bool DEG_SUM, int NT, int VT, int SEG_T>
__global__ void __launch_bounds__(NT) lbs_advance_kernel(LbsArgs a, Op op, int *__restrict__ out,
                                                         unsigned long long out_capacity,
                                                         unsigned long long *counters) {
    constexpr int NW = NT / 32;
    constexpr int WSTAGE = 32 * VT * 2;      // a tile adds at most 32*VT per warp
    __shared__ LbsSmem<NT, VT, SEG_T> sm;
    __shared__ int wstage[OUT_MODE == OUT_COMPACT ? NW : 1][OUT_MODE == OUT_COMPACT ? WSTAGE : 1];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    int *stage = wstage[OUT_MODE == OUT_COMPACT ? warp : 0];
    uint32_t wcnt = 0;                       // warp-uniform fill of this warp's stage
    unsigned long long deg_sum = 0;

    auto flush = [&]() {
        __syncwarp();
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(&counters[B200_CNT_OUT], (unsigned long long)wcnt);
        g = __shfl_sync(FULL_MASK, g, 0);
        for (uint32_t k = lane; k < wcnt; k += 32) {
            const int u = stage[k];
            if (g + k < out_capacity) out[g + k] = u;
            if (DEG_SUM) deg_sum += __ldg(a.offsets + u + 1) - __ldg(a.offsets + u);
        }
        if (lane == 0 && g + wcnt > out_capacity) counters[B200_CNT_OVERFLOW] = 1ull;
        wcnt = 0;
        __syncwarp();
    };

    lbs_for_each_tile<NT, VT, SEG_T>(a, sm, [&](uint32_t first_arc, uint32_t n_arcs, int j_lo, int j_hi, uint32_t) {
        int src[VT], dst[VT];
        uint32_t eid[VT], rank[VT];
        bool cand[VT];
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t k = threadIdx.x + i * NT;
            cand[i] = k < n_arcs;
            if (cand[i]) {
                const uint32_t arc = first_arc + k;
                const int j = lbs_locate(sm.start, j_lo, j_hi, arc);
                src[i] = sm.vert[j];
                eid[i] = sm.base[j] + arc;
                rank[i] = arc - sm.start[j];
                dst[i] = ld_stream(a.indices + eid[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < VT; ++i)
            if (cand[i] && OUT_MODE != OUT_RAW) cand[i] = op.probe(src[i], dst[i], eid[i]);
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t k = threadIdx.x + i * NT;
            bool emit = false;
            if (OUT_MODE == OUT_RAW) {
                if (cand[i]) {
                    emit = op.probe(src[i], dst[i], eid[i]) && op.commit(src[i], dst[i], eid[i], rank[i], first_arc + k);
                    if ((unsigned long long)first_arc + k < out_capacity) out[first_arc + k] = emit ? dst[i] : -1;
                }
            } else {
                if (cand[i]) emit = op.commit(src[i], dst[i], eid[i], rank[i], first_arc + k);
                if (OUT_MODE == OUT_COMPACT) {
                    const unsigned mask = __ballot_sync(FULL_MASK, emit);
                    if (emit) stage[wcnt + __popc(mask & lanemask_lt())] = dst[i];
                    wcnt += __popc(mask);
                }
            }
        }
        if (OUT_MODE == OUT_COMPACT && wcnt > WSTAGE - 32 * VT) flush();
    });
    if (OUT_MODE == OUT_COMPACT && wcnt) flush();
    if (DEG_SUM) {
        __shared__ unsigned long long s_deg;
        if (threadIdx.x == 0) s_deg = 0;
        __syncthreads();
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) deg_sum += __shfl_xor_sync(FULL_MASK, deg_sum, d);
        if (lane == 0 && deg_sum) atomicAdd(&s_deg, deg_sum);
        __syncthreads();
        if (threadIdx.x == 0 && s_deg) atomicAdd(&counters[B200_CNT_AUX], s_deg);
    }
    if (OUT_MODE == OUT_RAW && blockIdx.x == 0 && threadIdx.x == 0) {
        counters[B200_CNT_OUT] = *a.total;
        if (*a.total > out_capacity) counters[B200_CNT_OVERFLOW] = 1ull;
    }
}

// ---------------------------------------------------------------------------
// Built-in per-arc functors (the engine-side forms of bfs_functor_t /
// sssp_functor_t; the reference functors themselves run through
// RefFunctorOp in include/gunrock/advance.hxx).
// ---------------------------------------------------------------------------

// BFS push with a 1-bit visited map kept beside the labels:
//   cond_advance  (bfs_functor.hxx:26-28)  labels[dst] == -1   -> bit test
//   apply_advance (bfs_functor.hxx:30-33)  atomicCAS(-1 -> it+1) -> atomicOr wins exactly once
// The bitmap is n/8 bytes (512 KiB at scale 22, 8 MiB at scale 26) and stays in
// L2 / L1; the label itself is written once per discovered vertex.
// deg_sum (optional) accumulates the degree of every discovered vertex so the
// next level's direction can be chosen without another pass.
struct BfsPushOp {
    uint32_t *visited;
    int *labels;
    int next_label;
    __device__ __forceinline__ bool probe(int, int dst, uint32_t) const {
        return !((visited[dst >> 5] >> (dst & 31)) & 1u);
    }
    __device__ __forceinline__ bool commit(int, int dst, uint32_t, uint32_t, uint32_t) const {
        const uint32_t bit = 1u << (dst & 31);
        if (atomicOr(visited + (dst >> 5), bit) & bit) return false;
        labels[dst] = next_label;
        return true;
    }
};

// Idempotent BFS advance (advance.hxx:60 with idempotence=true: every neighbour is
// emitted, duplicates included); visited-bit candidates only, so the following
// uniquify filter sees far fewer items than m_F.
struct BfsIdempotentOp {
    const uint32_t *visited;
    __device__ __forceinline__ bool probe(int, int dst, uint32_t) const {
        return !((visited[dst >> 5] >> (dst & 31)) & 1u);
    }
    __device__ __forceinline__ bool commit(int, int, uint32_t, uint32_t, uint32_t) const { return true; }
};

// SSSP relax (sssp_functor.hxx:20-34).  Distances are non-negative floats, so the
// CAS loop of intrinsics.hxx:12-22 becomes one native integer atomicMin on the
// bit pattern.  The per-iteration `visited` stamp of cond_filter
// (sssp_functor.hxx:12-18) is folded in with an atomicExch, which makes the
// dedupe exact instead of racy.  preds (nullable) is written only by improving
// relaxations (the reference writes it unconditionally).
struct SsspRelaxOp {
    float *dist;
    const float *weights;
    int *preds;
    int *stamp;      // nullable => idempotent output (duplicates kept)
    int iteration;
    __device__ __forceinline__ bool probe(int src, int dst, uint32_t eid) const {
        return dist[src] + ld_stream(weights + eid) < dist[dst];
    }
    __device__ __forceinline__ bool commit(int src, int dst, uint32_t eid, uint32_t, uint32_t) const {
        const float nd = dist[src] + ld_stream(weights + eid);
        const float old = __int_as_float(atomicMin(reinterpret_cast<int *>(dist + dst), __float_as_int(nd)));
        if (!(nd < old)) return false;
        if (preds) preds[dst] = src;
        if (stamp) return atomicExch(stamp + dst, iteration) != iteration;
        return true;
    }
};

// ---------------------------------------------------------------------------
// Pull (bottom-up) BFS step over bitmaps: replaces gen_unvisited_kernel +
// sparse_to_dense_kernel + advance_backward_kernel + filter_kernel
// (advance.hxx:69-160, bfs_enactor.hxx:74-113).  One lane per vertex, one warp
// per 32-vertex bitmap word; a lane stops at its first in-neighbour found in the
// frontier bitmap (the reference scans every in-arc of every unvisited vertex
// each iteration).  The warp owns its word of next/visited, so
// those are plain stores.
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) bfs_pull_kernel(uint32_t n, const uint32_t *__restrict__ offsets,
                                                      const int *__restrict__ indices,
                                                      const uint32_t *__restrict__ frontier_bm,
                                                      uint32_t *__restrict__ next_bm, uint32_t *__restrict__ visited_bm,
                                                      int *__restrict__ labels, int next_label,
                                                      unsigned long long *counters) {
    const uint32_t num_words = (n + 31) >> 5;
    const uint32_t warps_total = (gridDim.x * NT) >> 5;
    const unsigned lane = lane_id();
    unsigned long long found_cnt = 0, inspected = 0, deg_found = 0;
    for (uint32_t word = (blockIdx.x * NT + threadIdx.x) >> 5; word < num_words; word += warps_total) {
        const uint32_t vis = visited_bm[word];
        const uint32_t v = (word << 5) + lane;
        bool found = false;
        if (v < n && !((vis >> lane) & 1u)) {
            const uint32_t b = __ldg(offsets + v), e = __ldg(offsets + v + 1);
            for (uint32_t k = b; k < e; ++k) {
                const int u = __ldg(indices + k);
                ++inspected;
                if ((__ldg(frontier_bm + (u >> 5)) >> (u & 31)) & 1u) { found = true; break; }
            }
            if (found) {
                labels[v] = next_label;
                deg_found += e - b;
            }
        }
        const unsigned mask = __ballot_sync(FULL_MASK, found);
        if (lane == 0) {
            next_bm[word] = mask;
            if (mask) visited_bm[word] = vis | mask;
            found_cnt += __popc(mask);
        }
    }
    // CTA reduction of the three counters -> 3 atomics per CTA
    __shared__ unsigned long long red[3][NT / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        inspected += __shfl_xor_sync(FULL_MASK, inspected, d);
        deg_found += __shfl_xor_sync(FULL_MASK, deg_found, d);
    }
    if (lane == 0) {
        red[0][threadIdx.x >> 5] = found_cnt;
        red[1][threadIdx.x >> 5] = inspected;
        red[2][threadIdx.x >> 5] = deg_found;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long s = 0;
        for (int w = 0; w < NT / 32; ++w) s += red[threadIdx.x][w];
        const int slot = threadIdx.x == 0 ? B200_CNT_OUT : (threadIdx.x == 1 ? B200_CNT_ARCS : B200_CNT_AUX);
        if (s) atomicAdd(&counters[slot], s);
    }
}

// frontier list -> bitmap (sparse_to_dense_kernel, advance.hxx:69-84); bitmap pre-cleared.
static __global__ void sparse_to_bitmap_kernel(const int *__restrict__ sparse, uint32_t len, uint32_t *bitmap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) {
        const int v = sparse[i];
        if (v >= 0) atomicOr(bitmap + (v >> 5), 1u << (v & 31));
    }
}

}  // namespace b200
