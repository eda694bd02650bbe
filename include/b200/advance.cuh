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
#include "loop_dyn.cuh"
#include "device_utils.cuh"
#include "lbs.cuh"
#include "workspace.h"
#include "partition.cuh"

namespace b200 {

enum { OUT_NONE = 0, OUT_COMPACT = 1, OUT_RAW = 2, OUT_ROUTED = 3 };

// OUT_ROUTED (multi-GPU push): every accepted vertex goes to the box of its owner
// (op.route(u) in [0, num_dest)); box[me] is this rank's next frontier, the others are
// the per-peer send buffers of the alltoallv that follows.  Bucketing happens inside the
// advance kernel's flush, so no separate partition pass runs before the exchange.
constexpr int MAX_DEST = 8;
struct RoutedOut {
    int *box[MAX_DEST];
    unsigned long long capacity[MAX_DEST];
    unsigned long long *count;   // [MAX_DEST] device counters
    int num_dest;
    int self;                    // graph-driven loop: box[self] is replaced by LoopDyn::out (the local next frontier)
};

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
                                                         unsigned long long *counters, RoutedOut routed) {
    constexpr int NW = NT / 32;
    constexpr bool STAGED = OUT_MODE == OUT_COMPACT || OUT_MODE == OUT_ROUTED;
    constexpr int WSTAGE = 32 * VT * 2;      // a tile adds at most 32*VT per warp
    __shared__ LbsSmem<NT, VT, SEG_T> sm;
    __shared__ int wstage[STAGED ? NW : 1][STAGED ? WSTAGE : 1];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    int *stage = wstage[STAGED ? warp : 0];
    uint32_t wcnt = 0;                       // warp-uniform fill of this warp's stage
    unsigned long long deg_sum = 0;

    auto flush = [&]() {
        __syncwarp();
        if constexpr (OUT_MODE == OUT_ROUTED) {
            // pass 1: how many staged vertices go to each destination (warp totals)
            uint32_t cnt[MAX_DEST];
#pragma unroll
            for (int p = 0; p < MAX_DEST; ++p) cnt[p] = 0;
            for (uint32_t k = lane; k < wcnt; k += 32) {
                const int d = op.route(stage[k]);
#pragma unroll
                for (int p = 0; p < MAX_DEST; ++p) cnt[p] += (d == p);
            }
            unsigned long long base[MAX_DEST];
#pragma unroll
            for (int p = 0; p < MAX_DEST; ++p) {
                base[p] = 0;
                if (p < routed.num_dest) {
                    const uint32_t c = warp_sum(cnt[p]);
                    if (c) {
                        if (lane == 0) base[p] = atomicAdd(&routed.count[p], (unsigned long long)c);
                        base[p] = __shfl_sync(FULL_MASK, base[p], 0);
                        if (lane == 0 && base[p] + c > routed.capacity[p]) counters[B200_CNT_OVERFLOW] = 1ull;
                    }
                }
            }
            // pass 2: scatter, 32 staged vertices at a time, keeping each destination's run contiguous
            for (uint32_t k0 = 0; k0 < wcnt; k0 += 32) {
                const uint32_t k = k0 + lane;
                const int u = k < wcnt ? stage[k] : -1;
                const int d = k < wcnt ? op.route(u) : -1;
#pragma unroll
                for (int p = 0; p < MAX_DEST; ++p) {
                    if (p < routed.num_dest) {
                        const unsigned mask = __ballot_sync(FULL_MASK, d == p);
                        if (d == p) {
                            const unsigned long long pos = base[p] + __popc(mask & lanemask_lt());
                            if (pos < routed.capacity[p]) routed.box[p][pos] = u;
                        }
                        base[p] += __popc(mask);
                    }
                }
                if (DEG_SUM && u >= 0 && op.route_is_local(d)) {
                    const uint32_t r = (uint32_t)u >> a.row_shift;
                    deg_sum += __ldg(a.offsets + r + 1) - __ldg(a.offsets + r);
                }
            }
            wcnt = 0;
            __syncwarp();
            return;
        }
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(&counters[B200_CNT_OUT], (unsigned long long)wcnt);
        g = __shfl_sync(FULL_MASK, g, 0);
        for (uint32_t k = lane; k < wcnt; k += 32) {
            const int u = stage[k];
            if (g + k < out_capacity) out[g + k] = u;
            if (DEG_SUM) {
                const uint32_t r = (uint32_t)u >> a.row_shift;
                deg_sum += __ldg(a.offsets + r + 1) - __ldg(a.offsets + r);
            }
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
                if (STAGED) {
                    const unsigned mask = __ballot_sync(FULL_MASK, emit);
                    if (emit) stage[wcnt + __popc(mask & lanemask_lt())] = dst[i];
                    wcnt += __popc(mask);
                }
            }
        }
        if (STAGED && wcnt > WSTAGE - 32 * VT) flush();
    });
    if (STAGED && wcnt) flush();
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

// BFS push on one rank's slice.  `known` is an n-bit map: exact "visited" for owned
// vertices, "already sent to its owner / known visited" for the others, so every remote
// vertex crosses NVLink at most once per sending rank over the whole traversal.
struct BfsPushPartOp {
    uint32_t *known;
    int *labels;         // local: labels[row(v)]
    int next_label;
    Partition part;
    __device__ __forceinline__ bool probe(int, int dst, uint32_t) const {
        const uint32_t b = part.bit((uint32_t)dst);
        return !((known[b >> 5] >> (b & 31)) & 1u);
    }
    __device__ __forceinline__ bool commit(int, int dst, uint32_t, uint32_t, uint32_t) const {
        const uint32_t b = part.bit((uint32_t)dst);
        const uint32_t bit = 1u << (b & 31);
        if (atomicOr(known + (b >> 5), bit) & bit) return false;
        if (part.owner((uint32_t)dst) == part.me) labels[part.row((uint32_t)dst)] = next_label;
        return true;
    }
    __device__ __forceinline__ int route(int u) const { return (int)part.owner((uint32_t)u); }
    __device__ __forceinline__ bool route_is_local(int d) const { return d == (int)part.me; }
};

// Receiver side of the push exchange: vertices owned by this rank that peers discovered.
// Test-and-set the visited bit; survivors get their label and join the next frontier
// (appended after the local discoveries: same counter).  This is the uniquify filter of
// the multi-GPU path.
static __global__ void bfs_absorb_kernel(const int *__restrict__ inbox, uint32_t count, uint32_t *known, int *labels,
                                         int next_label, Partition part, int *next_frontier,
                                         unsigned long long capacity, unsigned long long *next_count,
                                         unsigned long long *counters, const uint32_t *__restrict__ offsets) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool fresh = false;
    int u = -1;
    if (i < count) {
        u = inbox[i];
        const uint32_t b = part.bit((uint32_t)u);
        const uint32_t bit = 1u << (b & 31);
        fresh = !(known[b >> 5] & bit) && !(atomicOr(known + (b >> 5), bit) & bit);
        if (fresh) labels[part.row((uint32_t)u)] = next_label;
    }
    const unsigned mask = __ballot_sync(FULL_MASK, fresh);
    if (mask) {
        unsigned long long base = 0;
        const unsigned leader = __ffs(mask) - 1;
        if (lane_id() == leader) base = atomicAdd(next_count, (unsigned long long)__popc(mask));
        base = __shfl_sync(FULL_MASK, base, leader);
        unsigned long long deg = 0;
        if (fresh) {
            const unsigned long long pos = base + __popc(mask & lanemask_lt());
            if (pos < capacity) next_frontier[pos] = u;
            else counters[B200_CNT_OVERFLOW] = 1ull;
            const uint32_t r = part.row((uint32_t)u);
            deg = __ldg(offsets + r + 1) - __ldg(offsets + r);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) deg += __shfl_xor_sync(FULL_MASK, deg, d);
        if (lane_id() == leader && deg) atomicAdd(&counters[B200_CNT_AUX], deg);
    }
}

// dst[i] |= src[i]  (fold the all-gathered frontier bitmap into `known` after a pull level)
static __global__ void bitmap_or_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, size_t words) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) dst[i] |= src[i];
}

// ---------------------------------------------------------------------------
// Pull (bottom-up) BFS step over bitmaps: replaces gen_unvisited_kernel +
// sparse_to_dense_kernel + advance_backward_kernel + filter_kernel
// (advance.hxx:69-160, bfs_enactor.hxx:74-113).  One lane per vertex, one warp
// per 32-vertex bitmap word; a lane stops at its first in-neighbour found in the
// frontier bitmap (the reference scans every in-arc of every unvisited vertex
// each iteration).  The warp owns its word of next/visited, so
// those are plain stores.
// ---------------------------------------------------------------------------
// Partitioned form: the warp walks this rank's LOCAL rows (word w of its slice), `frontier_bm`
// is the whole (all-gathered) frontier bitmap, `next_bm` / `visited_bm` point at this rank's
// slice of the next-frontier / known bitmaps, labels are local.
#ifndef B200_PULL_MINB
#define B200_PULL_MINB 1
#endif
#ifndef B200_PULL_CW
#define B200_PULL_CW 32
#endif
// v2 (after the scale-26 level timings in profiles/README.md: the one-word-per-iteration walk had one load in
// flight per warp -- 0.12 ms to skip over an 8 MB bitmap whose vertices were all visited): a warp takes a CHUNK of
// B200_PULL_CW bitmap words with one coalesced load, skips it if nothing is unvisited, otherwise expands the unvisited bits
// into a dense per-warp list (shared memory, 16-bit ids) and walks the list U vertices per lane at a time, so the
// offsets of 4 vertices, then their first in-neighbours, then the 4 frontier-bitmap probes are in flight together.
// Most vertices stop at their first in-arc (scale-26 level 1: 1.14 arcs inspected per discovered vertex); the
// rest finish in a sequential early-exit loop.  Results and the inspected-arc count are the same as v1.
// COHERENT = true: the bitmaps are read through L2 (ld.global.cg).  For a kernel that runs SEVERAL pull levels in one
// launch (p2p_bfs.cu): the frontier / visited words are rewritten between its levels by other SMs, and neither L1 nor
// the non-coherent path is invalidated inside a launch.  The graph arrays never change: they stay on __ldg.
template <bool COHERENT>
__device__ __forceinline__ uint32_t pull_ld_bm(const uint32_t *p) {
    return COHERENT ? __ldcg(p) : __ldg(p);
}

// CW: bitmap words (32 CW vertices) per dynamically claimed warp chunk.  32 everywhere: 8-word chunks (tried for the
// ranks of a partition, which have < 2 chunks of 32 words per warp at scale 26 / 8 GPUs) made the sparse pull levels twice
// as slow -- their cost is the per-chunk claim and bitmap load, not the rows (single rank, scale 25: 57 -> 116 us).
template <int NT, bool COHERENT = false, int CW_ = B200_PULL_CW>
__device__ __forceinline__ void bfs_pull_body(uint32_t n, const uint32_t *__restrict__ offsets,
                                              const int *__restrict__ indices,
                                              const uint32_t *__restrict__ frontier_bm,
                                              uint32_t *__restrict__ next_bm, uint32_t *__restrict__ visited_bm,
                                              int *__restrict__ labels, int next_label,
                                              unsigned long long *counters, Partition part,
                                              const int *__restrict__ first_nbr = nullptr) {
    constexpr int NW = NT / 32, CW = CW_, U = 4;   // CW words (32 CW vertices) per warp chunk
    __shared__ uint16_t s_list[NW][CW * 32];
    __shared__ uint32_t s_new[NW][CW];
    __shared__ unsigned long long red[3][NW];
    const uint32_t num_words = (n + 31) >> 5;
    const uint32_t num_chunks = (num_words + CW - 1) / CW;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint16_t *list = s_list[warp];
    uint32_t *snew = s_new[warp];
    unsigned long long found_cnt = 0, inspected = 0, deg_found = 0;
    // chunks are claimed dynamically: on un-permuted RMAT the work of a chunk follows its id bits, and a strided
    // static deal gave every warp chunks of one kind (ncu: 27 % of the stall samples were finished warps waiting
    // at the CTA's last barrier)
    // (the claim of the NEXT chunk is issued before this one is processed, so its round trip is hidden)
    uint32_t pending = 0;
    if (lane == 0) pending = (uint32_t)atomicAdd(&counters[B200_CNT_WORK], 1ull);
    for (;;) {
        const uint32_t chunk = __shfl_sync(FULL_MASK, pending, 0);
        if (chunk >= num_chunks) break;
        if (lane == 0) pending = (uint32_t)atomicAdd(&counters[B200_CNT_WORK], 1ull);
        const uint32_t w = chunk * CW + lane;
        uint32_t vis = 0xffffffffu, unv = 0u;
        const bool have_w = lane < (unsigned)CW && w < num_words;
        if (have_w) {
            vis = COHERENT ? __ldcg(visited_bm + w) : visited_bm[w];
            unv = ~vis;
            const uint32_t rem = n - (w << 5);
            if (rem < 32u) unv &= (1u << rem) - 1u;      // bits past n in the last word
        }
        const uint32_t c = __popc(unv);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= (unsigned)d) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL_MASK, incl, 31);
        if (total == 0u) {                               // warp-uniform: nothing unvisited in these 1024 vertices
            if (have_w) next_bm[w] = 0u;
            continue;
        }
        if (lane < (unsigned)CW) snew[lane] = 0u;
        uint32_t pos = incl - c;
        while (unv) {
            const uint32_t b = __ffs(unv) - 1;
            unv &= unv - 1;
            list[pos++] = (uint16_t)((lane << 5) | b);
        }
        __syncwarp();
        const uint32_t v0 = chunk * (CW * 32);
        for (uint32_t k0 = 0; k0 < total; k0 += 32 * U) {
            uint32_t id[U], rb[U], re[U];
            bool on[U];
            int nb[U];
            if (first_nbr) {
                // derived array (b200_graph::first_in_neighbor): the first in-arc of every vertex, contiguous -- a
                // vertex that stops at its first arc (most do) touches neither the offsets nor the index array
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t k = k0 + u * 32 + lane;
                    on[u] = k < total;
                    id[u] = on[u] ? list[k] : 0u;
                    nb[u] = on[u] ? __ldg(first_nbr + v0 + id[u]) : -1;
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t k = k0 + u * 32 + lane;
                    on[u] = k < total;
                    id[u] = on[u] ? list[k] : 0u;
                    rb[u] = 0u;
                    re[u] = 0u;
                    if (on[u]) {
                        rb[u] = __ldg(offsets + v0 + id[u]);
                        re[u] = __ldg(offsets + v0 + id[u] + 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) nb[u] = re[u] > rb[u] ? __ldg(indices + rb[u]) : -1;
            }
            uint32_t hit[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                hit[u] = 0u;
                if (nb[u] >= 0) {
                    const uint32_t ub = part.bit((uint32_t)nb[u]);
                    hit[u] = (pull_ld_bm<COHERENT>(frontier_bm + (ub >> 5)) >> (ub & 31)) & 1u;
                }
            }
            if (first_nbr) {   // only the vertices that go on need their row bounds
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    rb[u] = 0u;
                    re[u] = 1u;
                    if (nb[u] >= 0 && !hit[u]) {
                        rb[u] = __ldg(offsets + v0 + id[u]);
                        re[u] = __ldg(offsets + v0 + id[u] + 1);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (nb[u] >= 0) {
                    bool found = hit[u] != 0u;
                    uint32_t insp = 1u;
                    // the rest of the row, B arcs per step: B index loads, then B probes, in flight together
                    // (the lanes of a warp wait for the longest row: fewer dependent steps, not fewer loads)
                    constexpr uint32_t B = 4;
                    for (uint32_t k = rb[u] + 1; !found && k < re[u]; k += B) {
                        int d[B];
#pragma unroll
                        for (uint32_t t = 0; t < B; ++t) d[t] = k + t < re[u] ? __ldg(indices + k + t) : -1;
                        uint32_t h = 0u;
#pragma unroll
                        for (uint32_t t = 0; t < B; ++t) {
                            if (d[t] >= 0) {
                                const uint32_t ub = part.bit((uint32_t)d[t]);
                                h |= ((pull_ld_bm<COHERENT>(frontier_bm + (ub >> 5)) >> (ub & 31)) & 1u) << t;
                            }
                        }
                        if (h) {
                            found = true;
                            insp += (uint32_t)__ffs(h);          // arcs up to and including the first parent
                        } else {
                            insp += re[u] - k < B ? re[u] - k : B;
                        }
                    }
                    inspected += insp;
                    if (found) {
                        labels[v0 + id[u]] = next_label;
                        if (!first_nbr) deg_found += re[u] - rb[u];   // (only push levels feed the direction decision)
                        atomicOr(&snew[id[u] >> 5], 1u << (id[u] & 31u));
                    }
                }
            }
        }
        __syncwarp();
        const uint32_t nbits = lane < (unsigned)CW ? snew[lane] : 0u;
        if (have_w) {
            next_bm[w] = nbits;
            if (nbits) visited_bm[w] = vis | nbits;
        }
        found_cnt += __popc(nbits);
        __syncwarp();                                    // the list and snew are reused by the next chunk
    }
    // CTA reduction of the three counters -> 3 atomics per CTA
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        found_cnt += __shfl_xor_sync(FULL_MASK, found_cnt, d);
        inspected += __shfl_xor_sync(FULL_MASK, inspected, d);
        deg_found += __shfl_xor_sync(FULL_MASK, deg_found, d);
    }
    if (lane == 0) {
        red[0][warp] = found_cnt;
        red[1][warp] = inspected;
        red[2][warp] = deg_found;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long s = 0;
        for (int w = 0; w < NW; ++w) s += red[threadIdx.x][w];
        const int slot = threadIdx.x == 0 ? B200_CNT_OUT : (threadIdx.x == 1 ? B200_CNT_ARCS : B200_CNT_AUX);
        if (s) atomicAdd(&counters[slot], s);
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, B200_PULL_MINB) bfs_pull_kernel(uint32_t n, const uint32_t *__restrict__ offsets,
                                                      const int *__restrict__ indices,
                                                      const uint32_t *__restrict__ frontier_bm,
                                                      uint32_t *__restrict__ next_bm, uint32_t *__restrict__ visited_bm,
                                                      int *__restrict__ labels, int next_label,
                                                      unsigned long long *counters, Partition part,
                                                      const int *__restrict__ first_nbr = nullptr) {
    bfs_pull_body<NT>(n, offsets, indices, frontier_bm, next_bm, visited_bm, labels, next_label, counters, part, first_nbr);
}

// Graph-driven level loop form: which bitmap is the frontier and the label come from the device (loop_dyn.cuh).
template <int NT>
__global__ void __launch_bounds__(NT, B200_PULL_MINB) bfs_pull_dyn_kernel(uint32_t n, const uint32_t *__restrict__ offsets,
                                                          const int *__restrict__ indices, uint32_t *bm0, uint32_t *bm1,
                                                          uint32_t *__restrict__ visited_bm, int *__restrict__ labels,
                                                          const LoopDyn *dyn, unsigned long long *counters, Partition part,
                                                          const int *__restrict__ first_nbr = nullptr) {
    if (!(dyn->run & LOOP_RUN_PULL)) return;
    const uint32_t bsel = dyn->bsel;
    bfs_pull_body<NT>(n, offsets, indices, bsel ? bm1 : bm0, bsel ? bm0 : bm1, visited_bm, labels, dyn->next_label,
                      counters, part, first_nbr);
}

// push -> pull switch of the graph-driven loop: frontier list -> bitmap (dynamic list and length; bitmap pre-cleared),
// and visited |= "vertex has no in-arc" (iso: prebuilt bitmap, else derived from the pull offsets; engine.cuh)
static __global__ void sparse_to_bitmap_dyn_kernel(const LoopDyn *dyn, uint32_t *bitmap, const uint32_t *__restrict__ pull_offsets,
                                                   uint32_t n, const uint32_t *__restrict__ iso, uint32_t *visited) {
    if (!(dyn->run & LOOP_RUN_TO_PULL)) return;
    or_no_in_arc_words(blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, pull_offsets, n, iso, visited);
    const int *__restrict__ sparse = dyn->in;
    const uint32_t len = dyn->len;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        const int v = sparse[i];
        if (v >= 0) atomicOr(bitmap + (v >> 5), 1u << (v & 31));
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
