// quad_advance.cuh -- second-generation push advance: merge-path LBS over ALIGNED QUADS.
//
// Same role as lbs_advance_kernel (advance.cuh), i.e. transform_scan + transform_lbs +
// the `!= -1` filter of the reference (advance.hxx:20-67, filter.hxx:11-31), redesigned
// after the ncu captures in profiles/ (the arc-wise kernel spent ~80 instructions per arc
// and sat at 30 % of the HBM roofline):
//
//  * the unit of load-balanced work is a 16-byte aligned QUAD of col_indices, not an
//    arc: a row [b, e) owns quads b/4 .. (e-1)/4, the degree scan counts quads, one
//    segment search serves four arcs, and every index (and weight) read is one aligned
//    128-bit ld.global.nc -- a warp reads 512 contiguous bytes per instruction.  The
//    arcs of a quad that fall outside [b, e) are masked, so results are unchanged;
//  * every WARP owns a contiguous chunk of the merged (quads U segment starts) list and
//    stages its own window of segments in its private slice of shared memory: there is
//    no CTA barrier anywhere in the walk;
//  * the per-arc functor runs in two stages (below): a cheap, fully unrolled probe of all
//    4*VT arcs a thread holds, then the survivors are compacted into a per-warp buffer and
//    the expensive part (atomics, label writes, output) runs densely, one candidate per lane;
//  * the visited bitmap is kept L1-RESIDENT: probes are ordinary (L1-allocating) loads while
//    the index/weight stream and the frontier arrays bypass L1 (ld.global.nc.L1::no_allocate),
//    and shared memory is kept small so the unified 256 KB SRAM is mostly cache.  A stale L1
//    word only costs a failed atomicOr; the global atomic decides who wins a vertex.
//    (Tried and rejected, numbers in profiles/README.md: (a) a software "hub" copy of the low-id
//    part of the bitmap in shared memory, updated or as a read-only pre-filter -- it takes the
//    SRAM away from L1 and was 1.7-2.4x slower on the heaviest scale-22 level; (b) re-dealing the
//    tile through shared memory so one probe instruction covers 32 consecutive neighbours -- 12 %
//    slower, most rows are too short for neighbouring arcs to share a bitmap line.)
//
// Stage 1 runs for every arc, unrolled so all loads of a thread are in flight together:
//   SrcVal   load_src(src)                          once per quad (e.g. dist[src])
//   Evidence probe_load(on, sv, src, dst, eid)      the load the test needs, predicated by `on`, branch-free
//   bool     probe_eval(ev, sv, dst, w)             false => reject
//   Cand     make_cand(sv, src, dst, eid, w)        what stage 2 needs (WORDS x uint32)
// Stage 2, dense, R candidates per lane in flight:
//   Token    claim (cand)                           the atomic; returns what it displaced
//   int      finish(token, cand)                    vertex to emit, or -1
#pragma once
#include "advance.cuh"
#include "loop_dyn.cuh"

namespace b200 {

struct QuadArgs {
    const int *frontier;               // [num_segments] vertex ids (-1 = hole)
    uint32_t num_segments;
    const uint32_t *scanned;           // [num_segments] exclusive scan of quads(v)
    const uint2 *rows;                 // [num_segments] (row begin, row end) of v, written by the scan
    const unsigned long long *total;   // device: Q = sum of quads(v)
    const uint32_t *offsets;           // row offsets [n+1] (degree of emitted vertices, DEG_SUM)
    const int4 *indices4;              // col_indices viewed as aligned quads
    const float4 *weights4;            // col_values viewed as aligned quads (weighted ops only)
    uint32_t min_chunk;                // lower bound on the work items per warp
    uint32_t row_shift;                // cyclic 1D partition: row of v is v >> row_shift
    const LoopDyn *dyn;                // graph-driven level loop: frontier / num_segments / out come from here (else NULL)
    // work creation (OUT_COMPACT only, else NULL): the flush also writes the NEXT level's row bounds and quad scan;
    // counters[B200_CNT_OUT] then holds (vertices emitted << 32) | quads created
    uint32_t *scanned_next;
    uint2 *rows_next;
};

// quads(v): number of aligned 16-byte quads of col_indices that row v touches.  Also leaves
// the row bounds next to the scan so the advance never chases frontier -> offsets itself.
struct FrontierQuads {
    const int *frontier;
    const uint32_t *offsets;
    uint2 *rows;
    uint32_t row_shift;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const int v = frontier[i];
        uint32_t b = 0, e = 0;
        if (v >= 0) {
            const uint32_t r = (uint32_t)v >> row_shift;
            b = __ldg(offsets + r);
            e = __ldg(offsets + r + 1);
        }
        rows[i] = make_uint2(b, e);
        return e > b ? ((e - 1) >> 2) - (b >> 2) + 1u : 0u;
    }
};

#ifndef B200_QUAD_SPLIT
#define B200_QUAD_SPLIT 8
#endif
#ifndef B200_QUAD_R
#define B200_QUAD_R 2
#endif
#ifndef B200_QUAD_MINB
#define B200_QUAD_MINB 1
#endif
constexpr int QUAD_R = B200_QUAD_R;          // stage-2 candidates per lane in flight
// per-warp shared-memory words: segment window + output stage + candidate buffer
template <class Op, bool STAGED, int VT, int WSEG, int WSTAGE>
__host__ __device__ constexpr int quad_warp_words() {
    return 4 * WSEG + (STAGED ? WSTAGE : 0) + (128 * VT + 32 * QUAD_R) * Op::Cand::WORDS;
}

template <class Op, int OUT_MODE, bool DEG_SUM, int NT, int VT, int WSEG>
__global__ void __launch_bounds__(NT, B200_QUAD_MINB)
quad_advance_kernel(QuadArgs a, Op op, int *__restrict__ out, unsigned long long out_capacity,
                    unsigned long long *counters, RoutedOut routed) {
    static_assert(OUT_MODE != OUT_RAW, "the raw (un-compacted) layout is produced by lbs_advance_kernel");
    using Cand = typename Op::Cand;
    constexpr int NW = NT / 32;
    constexpr int NA = 4 * VT;               // arcs per thread per tile
    constexpr int R = QUAD_R;
    constexpr int CCAP = 128 * VT + 32 * R;  // a tile adds <= 128*VT; < 32*R are left after a drain
    constexpr bool STAGED = OUT_MODE == OUT_COMPACT || OUT_MODE == OUT_ROUTED;
    constexpr int WSTAGE = 64 * R;           // a drain round adds at most 32*R
    // dynamic shared memory: one private area per warp
    extern __shared__ uint4 dyn_smem[];
    __shared__ unsigned long long s_sum[2];

    const unsigned lane = lane_id(), warp = threadIdx.x >> 5, lt_mask = lanemask_lt();
    if (a.dyn) {                             // level arguments decided on the device (loop_dyn.cuh)
        loop_trace(a.dyn, 3);
        if (!(a.dyn->run & LOOP_RUN_PUSH)) return;
        a.frontier = a.dyn->in;
        a.num_segments = a.dyn->len;
        out = a.dyn->out;
        if (a.dyn->scanned_in) {
            a.scanned = a.dyn->scanned_in;
            a.rows = a.dyn->rows_in;
            a.scanned_next = a.dyn->scanned_out;
            a.rows_next = a.dyn->rows_out;
        }
        if constexpr (OUT_MODE == OUT_ROUTED) {
#pragma unroll
            for (int p = 0; p < MAX_DEST; ++p)
                if (p == routed.self) routed.box[p] = a.dyn->out;
        }
    }
    const unsigned long long Q = *a.total;
    if (threadIdx.x < 2) s_sum[threadIdx.x] = 0;
    __syncthreads();

    uint32_t *warea = reinterpret_cast<uint32_t *>(dyn_smem) + warp * quad_warp_words<Op, STAGED, VT, WSEG, WSTAGE>();
    uint32_t *start = warea;                 // scanned quad start of staged segment j
    uint32_t *rb = warea + WSEG;             // row begin (arc id)
    uint32_t *re = warea + 2 * WSEG;         // row end
    int *vert = reinterpret_cast<int *>(warea + 3 * WSEG);
    int *stage = reinterpret_cast<int *>(warea + 4 * WSEG);
    uint32_t *cbuf = warea + 4 * WSEG + (STAGED ? WSTAGE : 0);   // [Cand::WORDS][CCAP]
    uint32_t wcnt = 0;                       // warp-uniform fill of this warp's output stage
    uint32_t ccnt = 0;                       // warp-uniform fill of the candidate buffer
    unsigned long long deg_sum = 0, arc_cnt = 0;

    auto flush = [&]() {
        __syncwarp();
        if constexpr (OUT_MODE == OUT_ROUTED) {
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
            for (uint32_t k0 = 0; k0 < wcnt; k0 += 32) {
                const uint32_t k = k0 + lane;
                const int u = k < wcnt ? stage[k] : -1;
                const int d = k < wcnt ? op.route(u) : -1;
#pragma unroll
                for (int p = 0; p < MAX_DEST; ++p) {
                    if (p < routed.num_dest) {
                        const unsigned mask = __ballot_sync(FULL_MASK, d == p);
                        if (d == p) {
                            const unsigned long long pos = base[p] + __popc(mask & lt_mask);
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
        } else if (a.scanned_next) {
            // work creation: one packed atomic reserves wcnt frontier slots AND the quads of their rows, so slot
            // order and scan order agree (the next level's merge-path search needs a sorted scan)
            // (one pass: the row bounds of the <= WSTAGE / 32 entries a lane owns stay in registers, all loads in flight)
            constexpr int E = WSTAGE / 32;
            uint32_t rb_[E], re_[E];
            uint32_t myq = 0;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const uint32_t k = lane + 32u * i;
                rb_[i] = 0u;
                re_[i] = 0u;
                if (k < wcnt) {
                    const uint32_t r = (uint32_t)stage[k] >> a.row_shift;
                    // (L1-bypassing loads here -- to keep the bitmap's L1 lines -- measured no gain: 0.2486 vs 0.2432 ms)
                    rb_[i] = __ldg(a.offsets + r);
                    re_[i] = __ldg(a.offsets + r + 1);
                }
            }
#pragma unroll
            for (int i = 0; i < E; ++i) {
                myq += re_[i] > rb_[i] ? ((re_[i] - 1) >> 2) - (rb_[i] >> 2) + 1u : 0u;
                if (DEG_SUM) deg_sum += re_[i] - rb_[i];
            }
            const uint32_t wq = warp_sum(myq);
            unsigned long long g64 = 0;
            if (lane == 0) g64 = atomicAdd(&counters[B200_CNT_OUT], ((unsigned long long)wcnt << 32) | wq);
            g64 = __shfl_sync(FULL_MASK, g64, 0);
            const unsigned long long g = g64 >> 32;
            uint32_t run = (uint32_t)g64;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const uint32_t k = lane + 32u * i;
                if (32u * i >= wcnt) break;                      // warp-uniform
                const bool on = k < wcnt;
                const uint32_t q = re_[i] > rb_[i] ? ((re_[i] - 1) >> 2) - (rb_[i] >> 2) + 1u : 0u;
                uint32_t incl = q;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL_MASK, incl, s);
                    if (lane >= (unsigned)s) incl += t;
                }
                if (on && g + k < out_capacity) {
                    out[g + k] = stage[k];
                    a.rows_next[g + k] = make_uint2(rb_[i], re_[i]);
                    a.scanned_next[g + k] = run + incl - q;
                }
                run += __shfl_sync(FULL_MASK, incl, 31);
            }
            if (lane == 0 && g + wcnt > out_capacity) counters[B200_CNT_OVERFLOW] = 1ull;
        } else {
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
        }
        wcnt = 0;
        __syncwarp();
    };

    // ---- stage 2: pop candidates off the end of the buffer, 32*R at a time, one per lane
    auto drain = [&](uint32_t leave) {       // until at most `leave` candidates remain
        __syncwarp();
        while (ccnt > leave) {
            const uint32_t n = ccnt < 32u * R ? ccnt : 32u * R, base = ccnt - n;
            Cand c[R];
            bool on[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t idx = base + 32u * r + lane;
                on[r] = idx < ccnt;
#pragma unroll
                for (int w = 0; w < Cand::WORDS; ++w) c[r].w[w] = on[r] ? cbuf[w * CCAP + idx] : 0u;
            }
            typename Op::Token tok[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                tok[r] = typename Op::Token();
                if (on[r]) tok[r] = op.claim(c[r]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int u = -1;
                if (on[r]) u = op.finish(tok[r], c[r]);
                if (STAGED) {
                    const unsigned mask = __ballot_sync(FULL_MASK, u >= 0);
                    if (u >= 0) stage[wcnt + __popc(mask & lt_mask)] = u;
                    wcnt += __popc(mask);
                }
            }
            if (STAGED && wcnt > WSTAGE - 32 * R) flush();
            ccnt = base;
        }
        __syncwarp();
    };

    // ---- stage 1 is software-pipelined over tiles: fetch(t+1) -- segment search and the 128-bit
    // index / weight loads -- is issued BEFORE probe(t), so the DRAM latency of the index stream
    // overlaps the L1/L2 latency of the previous tile's probes.  A tile's state lives in registers.
    struct TileRegs {
        int dst[NA];
        float wgt[Op::WEIGHTED ? NA : 1];
        uint32_t e0[VT];
        int src[VT];
        uint32_t valid;                      // bit 4i+t: arc t of quad i lies inside its row
    };
    // one tile = up to 32*VT quads, lane-strided so a warp reads 512 contiguous bytes per instruction
    auto fetch = [&](TileRegs &t, uint32_t first_q, uint32_t nq, int j_lo, int j_hi) {
        t.valid = 0u;
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t k = lane + 32u * i;
            t.e0[i] = 0u;
            t.src[i] = 0;
            int4 d = make_int4(0, 0, 0, 0);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < nq) {
                const uint32_t qa = first_q + k;
                const int j = lbs_locate(start, j_lo, j_hi, qa);
                const uint32_t b = rb[j], e = re[j];
                const uint32_t q = (b >> 2) + (qa - start[j]);
                d = ld_stream_v4(a.indices4 + q);
                if constexpr (Op::WEIGHTED) w = ld_stream_v4(a.weights4 + q);
                t.e0[i] = q << 2;
                const uint32_t lo = b > t.e0[i] ? b - t.e0[i] : 0u;
                const uint32_t hi = e - t.e0[i] < 4u ? e - t.e0[i] : 4u;
                t.valid |= (((1u << hi) - 1u) & ~((1u << lo) - 1u)) << (4 * i);
                t.src[i] = vert[j];
            }
            t.dst[4 * i] = d.x; t.dst[4 * i + 1] = d.y; t.dst[4 * i + 2] = d.z; t.dst[4 * i + 3] = d.w;
            if constexpr (Op::WEIGHTED) { t.wgt[4 * i] = w.x; t.wgt[4 * i + 1] = w.y; t.wgt[4 * i + 2] = w.z; t.wgt[4 * i + 3] = w.w; }
        }
        arc_cnt += __popc(t.valid);
    };
    auto probe = [&](const TileRegs &t) {
        typename Op::SrcVal sv[VT];
#pragma unroll
        for (int i = 0; i < VT; ++i) sv[i] = ((t.valid >> (4 * i)) & 15u) ? op.load_src(t.src[i]) : typename Op::SrcVal();
        // probe every arc: issue all loads (one L1-cached load each), then judge
        uint32_t cm = 0u;                    // candidate mask
        {
            typename Op::Evidence ev[NA];
#pragma unroll
            for (int k = 0; k < NA; ++k)
                ev[k] = op.probe_load((t.valid >> k) & 1u, sv[k >> 2], t.src[k >> 2], t.dst[k], t.e0[k >> 2] + (k & 3));
#pragma unroll
            for (int k = 0; k < NA; ++k)
                if (((t.valid >> k) & 1u) && op.probe_eval(ev[k], sv[k >> 2], t.dst[k], Op::WEIGHTED ? t.wgt[Op::WEIGHTED ? k : 0] : 0.f))
                    cm |= 1u << k;
        }
        // compact the survivors into the warp's candidate buffer
        if (__any_sync(FULL_MASK, cm != 0u)) {
            const uint32_t c = __popc(cm);
            uint32_t incl = c;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const uint32_t u = __shfl_up_sync(FULL_MASK, incl, s);
                if (lane >= (unsigned)s) incl += u;
            }
            const uint32_t pos0 = ccnt + incl - c;
#pragma unroll
            for (int k = 0; k < NA; ++k) {
                if ((cm >> k) & 1u) {
                    const Cand cd = op.make_cand(sv[k >> 2], t.src[k >> 2], t.dst[k], t.e0[k >> 2] + (k & 3), Op::WEIGHTED ? t.wgt[Op::WEIGHTED ? k : 0] : 0.f);
                    const uint32_t pos = pos0 + __popc(cm & ((1u << k) - 1u));
#pragma unroll
                    for (int w = 0; w < Cand::WORDS; ++w) cbuf[w * CCAP + pos] = cd.w[w];
                }
            }
            ccnt += __shfl_sync(FULL_MASK, incl, 31);
            if (ccnt >= 32u * R) drain(32u * R - 1u);
        }
    };
    TileRegs pend;                           // the tile whose loads are in flight
    bool have_pend = false;
    auto tile = [&](uint32_t first_q, uint32_t nq, int j_lo, int j_hi) {
        TileRegs next;
        fetch(next, first_q, nq, j_lo, j_hi);
        if (have_pend) probe(pend);
        pend = next;
        have_pend = true;
    };

    // ---- this warp's chunk of the merged (quads U segment starts) list.  Consecutive chunks go
    // to different SMs, so a level too small for the whole grid still spreads over the chip.
    const unsigned long long work = Q + a.num_segments;
    // Each warp takes SPLIT pieces, total_warps apart, so rows of very different cost (a hub row
    // whose sorted neighbours share bitmap lines vs. many short rows) average out over the warps.
    // The 2*SPLIT merge-path searches run in parallel lanes: one search latency in total.
    constexpr int SPLIT = B200_QUAD_SPLIT;
    static_assert(SPLIT >= 1 && SPLIT <= 16, "one lane per chunk boundary");
    const uint32_t total_warps = gridDim.x * NW, gw = warp * gridDim.x + blockIdx.x;
    unsigned long long chunk = ceil_div<unsigned long long>(work, (unsigned long long)total_warps * SPLIT);
    if (chunk < a.min_chunk) chunk = a.min_chunk;
    uint32_t sb = 0;
    {
        const unsigned long long piece = (unsigned long long)(lane >> 1) * total_warps + gw;
        unsigned long long d = (piece + (lane & 1u)) * chunk;
        if (d > work) d = work;
        if (Q != 0 && lane < 2u * SPLIT) sb = merge_path_segments(a.scanned, a.num_segments, Q, d);
    }
    for (int c = 0; c < SPLIT; ++c) {
        const unsigned long long d0 = ((unsigned long long)c * total_warps + gw) * chunk;
        const uint32_t s0 = __shfl_sync(FULL_MASK, sb, 2 * c), s1 = __shfl_sync(FULL_MASK, sb, 2 * c + 1);
        if (Q == 0 || d0 >= work) break;
        const unsigned long long d1 = d0 + chunk < work ? d0 + chunk : work;
        const uint32_t q0 = (uint32_t)(d0 - s0), q1 = (uint32_t)(d1 - s1);
        uint32_t cur_s = s0 > 0 ? s0 - 1 : 0;   // segment that contains quad q0 (or an empty one just before it)
        uint32_t cur_q = q0;
        while (cur_q < q1) {                     // one iteration per window of <= WSEG segments
            const int ns = (int)min((uint32_t)WSEG, s1 - cur_s);
            for (int j = lane; j < ns; j += 32) {
                const uint2 row = ld_stream_v2(a.rows + cur_s + j);
                start[j] = ld_stream(a.scanned + cur_s + j);
                vert[j] = ld_stream(a.frontier + cur_s + j);
                rb[j] = row.x;
                re[j] = row.y;
            }
            __syncwarp();
            uint32_t win_end = q1;
            bool more = false;
            if (cur_s + (uint32_t)ns < s1) {
                const uint32_t lim = ld_stream(a.scanned + cur_s + ns);
                if (lim < win_end) { win_end = lim; more = true; }
            }
            int j_lo = 0;
            while (cur_q < win_end) {
                const uint32_t q_end = win_end - cur_q > 32u * VT ? cur_q + 32u * VT : win_end;
                // (tiles of a long row: the next staged segment starts beyond the tile, nothing to search)
                if (j_lo + 1 < ns && start[j_lo + 1] <= cur_q) j_lo = lbs_locate(start, j_lo + 1, ns, cur_q);
                int j_hi = j_lo + 1;
                if (j_hi < ns && start[j_hi] <= q_end - 1) j_hi = lbs_locate(start, j_hi, ns, q_end - 1) + 1;
                tile(cur_q, q_end - cur_q, j_lo, j_hi);
                cur_q = q_end;
            }
            if (!more) break;                    // win_end == q1: piece done
            cur_s += (uint32_t)ns;               // scanned[cur_s + ns] == win_end == cur_q
            __syncwarp();
        }
        __syncwarp();
    }
    if (have_pend) probe(pend);
    if (ccnt) drain(0u);
    if (STAGED && wcnt) flush();

    // ---- CTA totals: arcs walked (m_F) and, optionally, the degree sum of the emitted vertices
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        arc_cnt += __shfl_xor_sync(FULL_MASK, arc_cnt, s);
        if (DEG_SUM) deg_sum += __shfl_xor_sync(FULL_MASK, deg_sum, s);
    }
    if (lane == 0) {
        if (arc_cnt) atomicAdd(&s_sum[0], arc_cnt);
        if (DEG_SUM && deg_sum) atomicAdd(&s_sum[1], deg_sum);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_sum[0]) atomicAdd(&counters[B200_CNT_ARCS], s_sum[0]);
        if (DEG_SUM && s_sum[1]) atomicAdd(&counters[B200_CNT_AUX], s_sum[1]);
    }
}

// ---------------------------------------------------------------------------
// Ops
// ---------------------------------------------------------------------------
struct NoSrc {};
template <int N> struct CandWords {
    static constexpr int WORDS = N;
    uint32_t w[N];
};

// Label stores are scattered 4-byte writes nobody in this launch reads back: tuning builds can keep them out of L1
// (-DB200_LABEL_STCG), which belongs to the visited bitmap.
__device__ __forceinline__ void store_label(int *p, int v) {
#ifdef B200_LABEL_STCG
    __stcg(p, v);
#else
    *p = v;
#endif
}

// BFS push (bfs_functor.hxx:26-33): cond_advance = bit test (L1-resident bitmap word),
// apply_advance = atomicOr wins exactly once.
struct BfsPushQ {
    uint32_t *visited;
    int *labels;
    int next_label;
    static constexpr bool WEIGHTED = false;
    using SrcVal = NoSrc;
    using Token = uint32_t;
    using Evidence = uint32_t;
    using Cand = CandWords<1>;   // dst
    __device__ __forceinline__ SrcVal load_src(int) const { return NoSrc(); }
    __device__ __forceinline__ Evidence probe_load(bool on, SrcVal, int, int dst, uint32_t) const {
        uint32_t word = 0xffffffffu;
#ifdef B200_PROBE_EVICT_LAST
        if (on) asm volatile("ld.global.L1::evict_last.u32 %0, [%1];" : "=r"(word) : "l"(visited + ((uint32_t)dst >> 5)));
#else
        if (on) word = visited[(uint32_t)dst >> 5];
#endif
        return word;
    }
    __device__ __forceinline__ bool probe_eval(Evidence word, SrcVal, int dst, float) const {
        return !((word >> (dst & 31)) & 1u);
    }
    __device__ __forceinline__ Cand make_cand(SrcVal, int, int dst, uint32_t, float) const { return Cand{{(uint32_t)dst}}; }
    __device__ __forceinline__ Token claim(const Cand &c) const {
        return atomicOr(visited + (c.w[0] >> 5), 1u << (c.w[0] & 31));
    }
    __device__ __forceinline__ int finish(Token old, const Cand &c) const {
        if ((old >> (c.w[0] & 31)) & 1u) return -1;
        store_label(labels + c.w[0], next_label);
        return (int)c.w[0];
    }
};

// BfsPushQ whose label comes from the device-resident level state (graph-driven loop).
struct BfsPushQDyn : BfsPushQ {
    const LoopDyn *dyn;
    __device__ __forceinline__ int finish(Token old, const Cand &c) const {
        if ((old >> (c.w[0] & 31)) & 1u) return -1;
        store_label(labels + c.w[0], dyn->next_label);
        return (int)c.w[0];
    }
};

// Idempotent BFS advance (advance.hxx:60, idempotence = true): every not-yet-visited
// neighbour is emitted, duplicates included; b200_uniquify follows.
struct BfsIdempotentQ {
    const uint32_t *visited;
    static constexpr bool WEIGHTED = false;
    using SrcVal = NoSrc;
    using Token = NoSrc;
    using Evidence = uint32_t;
    using Cand = CandWords<1>;
    __device__ __forceinline__ SrcVal load_src(int) const { return NoSrc(); }
    __device__ __forceinline__ Evidence probe_load(bool on, SrcVal, int, int dst, uint32_t) const {
        uint32_t word = 0xffffffffu;
        if (on) word = visited[(uint32_t)dst >> 5];
        return word;
    }
    __device__ __forceinline__ bool probe_eval(Evidence word, SrcVal, int dst, float) const {
        return !((word >> (dst & 31)) & 1u);
    }
    __device__ __forceinline__ Cand make_cand(SrcVal, int, int dst, uint32_t, float) const { return Cand{{(uint32_t)dst}}; }
    __device__ __forceinline__ Token claim(const Cand &) const { return NoSrc(); }
    __device__ __forceinline__ int finish(Token, const Cand &c) const { return (int)c.w[0]; }
};

// BFS push on one rank's slice of a cyclic 1D partition (see BfsPushPartOp): `known` is the
// n-bit rank-major map, labels are local.
struct BfsPushPartQ {
    uint32_t *known;
    int *labels;
    int next_label;
    Partition part;
    static constexpr bool WEIGHTED = false;
    using SrcVal = NoSrc;
    using Token = uint32_t;
    using Evidence = uint32_t;
    using Cand = CandWords<1>;
    __device__ __forceinline__ SrcVal load_src(int) const { return NoSrc(); }
    __device__ __forceinline__ Evidence probe_load(bool on, SrcVal, int, int dst, uint32_t) const {
        uint32_t word = 0xffffffffu;
        if (on) word = known[part.bit((uint32_t)dst) >> 5];
        return word;
    }
    __device__ __forceinline__ bool probe_eval(Evidence word, SrcVal, int dst, float) const {
        return !((word >> (part.bit((uint32_t)dst) & 31)) & 1u);
    }
    __device__ __forceinline__ Cand make_cand(SrcVal, int, int dst, uint32_t, float) const { return Cand{{(uint32_t)dst}}; }
    __device__ __forceinline__ Token claim(const Cand &c) const {
        const uint32_t b = part.bit(c.w[0]);
        return atomicOr(known + (b >> 5), 1u << (b & 31));
    }
    __device__ __forceinline__ int finish(Token old, const Cand &c) const {
        if ((old >> (part.bit(c.w[0]) & 31)) & 1u) return -1;
        if (part.owner(c.w[0]) == part.me) labels[part.row(c.w[0])] = next_label;
        return (int)c.w[0];
    }
    __device__ __forceinline__ int route(int u) const { return (int)part.owner((uint32_t)u); }
    __device__ __forceinline__ bool route_is_local(int d) const { return d == (int)part.me; }
};

// Multi-GPU push level whose exchange is the `known` bitmap itself (p2p_bfs.cu, big levels): the advance only CLAIMS
// the bit of every unvisited neighbour -- owner or not -- and emits nothing; after the level barrier each owner ORs
// its slice of every rank's `known` and labels what is new.  The atomicOr result is unused, so it compiles to a
// fire-and-forget RED.
struct BfsClaimPartQ {
    uint32_t *known;
    Partition part;
    static constexpr bool WEIGHTED = false;
    using SrcVal = NoSrc;
    using Token = NoSrc;
    using Evidence = uint32_t;
    using Cand = CandWords<1>;   // bit index in the rank-major map
    __device__ __forceinline__ SrcVal load_src(int) const { return NoSrc(); }
    __device__ __forceinline__ Evidence probe_load(bool on, SrcVal, int, int dst, uint32_t) const {
        uint32_t word = 0xffffffffu;
        if (on) word = known[part.bit((uint32_t)dst) >> 5];
        return word;
    }
    __device__ __forceinline__ bool probe_eval(Evidence word, SrcVal, int dst, float) const {
        return !((word >> (part.bit((uint32_t)dst) & 31)) & 1u);
    }
    __device__ __forceinline__ Cand make_cand(SrcVal, int, int dst, uint32_t, float) const {
        return Cand{{part.bit((uint32_t)dst)}};
    }
    __device__ __forceinline__ Token claim(const Cand &c) const {
        atomicOr(known + (c.w[0] >> 5), 1u << (c.w[0] & 31));
        return NoSrc();
    }
    __device__ __forceinline__ int finish(Token, const Cand &) const { return -1; }
};

// BfsPushPartQ whose label comes from the device-resident level state (graph-driven loop).
struct BfsPushPartQDyn : BfsPushPartQ {
    const LoopDyn *dyn;
    __device__ __forceinline__ int finish(Token old, const Cand &c) const {
        if ((old >> (part.bit(c.w[0]) & 31)) & 1u) return -1;
        if (part.owner(c.w[0]) == part.me) labels[part.row(c.w[0])] = dyn->next_label;
        return (int)c.w[0];
    }
};

// SSSP relax (sssp_functor.hxx:20-34), see SsspRelaxOp.  dist[src] is read once per quad.
struct SsspRelaxQ {
    float *dist;
    int *preds;
    int *stamp;      // nullable => idempotent output (duplicates kept)
    int iteration;
    const NearFar *nf;   // nullable => every improved vertex is emitted (the reference's Bellman-Ford iterations)
    static constexpr bool WEIGHTED = true;
    using SrcVal = float;
    using Token = float;
    using Evidence = float;
    using Cand = CandWords<3>;   // dst, bits of dist[src] + w, src
    __device__ __forceinline__ SrcVal load_src(int src) const { return __ldcg(dist + src); }
    __device__ __forceinline__ Evidence probe_load(bool on, SrcVal, int, int dst, uint32_t) const {
        float d = 0.f;
        if (on) d = dist[dst];
        return d;
    }
    __device__ __forceinline__ bool probe_eval(Evidence dd, SrcVal ds, int, float w) const { return ds + w < dd; }
    __device__ __forceinline__ Cand make_cand(SrcVal ds, int src, int dst, uint32_t, float w) const {
        return Cand{{(uint32_t)dst, (uint32_t)__float_as_int(ds + w), (uint32_t)src}};
    }
    __device__ __forceinline__ Token claim(const Cand &c) const {
        return __int_as_float(atomicMin(reinterpret_cast<int *>(dist + c.w[0]), (int)c.w[1]));
    }
    __device__ __forceinline__ int finish(Token old, const Cand &c) const {
        if (!(__int_as_float((int)c.w[1]) < old)) return -1;
        if (preds) preds[c.w[0]] = (int)c.w[2];
        if (nf && !(__int_as_float((int)c.w[1]) < nf->cutoff)) return -1;   // pending: its bucket will take it
        if (stamp && atomicExch(stamp + c.w[0], iteration) == iteration) return -1;
        return (int)c.w[0];
    }
};

// SsspRelaxQ whose iteration stamp comes from the device-resident level state (graph-driven loop).
struct SsspRelaxQDyn : SsspRelaxQ {
    const LoopDyn *dyn;
    __device__ __forceinline__ int finish(Token old, const Cand &c) const {
        if (!(__int_as_float((int)c.w[1]) < old)) return -1;
        if (preds) preds[c.w[0]] = (int)c.w[2];
        if (nf && !(__int_as_float((int)c.w[1]) < nf->cutoff)) return -1;
        const int it = dyn->next_label - 1;
        if (stamp && atomicExch(stamp + c.w[0], it) == it) return -1;
        return (int)c.w[0];
    }
};

// Deterministic predecessors once the distances are final: the smallest u != v with
// dist[u] + w(u,v) == dist[v] (the reference's GPU preds are a race).  A zero-weight
// arc between two vertices of equal distance counts only from the smaller id to the larger one, so
// parent chains never close a cycle (a zero-weight self loop or u <-> v pair cannot make preds[v] = v).
struct SsspPredQ {
    const float *dist;
    int *preds;
    static constexpr bool WEIGHTED = true;
    using SrcVal = float;
    using Token = NoSrc;
    using Evidence = float;
    using Cand = CandWords<2>;   // dst, src
    __device__ __forceinline__ SrcVal load_src(int src) const { return dist[src]; }
    __device__ __forceinline__ Evidence probe_load(bool on, SrcVal, int, int dst, uint32_t) const {
        float d = 0.f;
        if (on) d = dist[dst];
        return d;
    }
    __device__ __forceinline__ bool probe_eval(Evidence dd, SrcVal ds, int, float w) const {
        return ds != 3.402823466e+38f && ds + w == dd;
    }
    __device__ __forceinline__ Cand make_cand(SrcVal ds, int src, int dst, uint32_t, float w) const {
        // (the tie rule needs src, which probe_eval does not see: an excluded pair becomes a no-op candidate)
        const bool ok = src != dst && (w > 0.f || src < dst);
        (void)ds;
        return Cand{{(uint32_t)dst, ok ? (uint32_t)src : 0x7fffffffu}};
    }
    __device__ __forceinline__ Token claim(const Cand &c) const {
        atomicMin(preds + c.w[0], (int)c.w[1]);
        return NoSrc();
    }
    __device__ __forceinline__ int finish(Token, const Cand &) const { return -1; }
};

}  // namespace b200
