// quad_segreduce.cuh -- second-generation neighborhood reduce: the quad walk of quad_advance.cuh
// with a warp-segmented reduction instead of the probe / claim stages.
//
// Same role as lbs_segreduce_kernel (segreduce.cuh), i.e. mgpu lbs_segreduce as used by
// neighborhood_kernel (neighborhood.hxx:47-58; kernel_segreduce.hxx:306-387), redesigned after
// the ncu capture of the arc-wise kernel (profiles/r01_ncu_segreduce_v1_lbs.txt: 128 thread
// instructions per arc, 48 % of the stall samples on the index -> value dependency, 15 % of the
// HBM roofline on the full scale-24 frontier):
//
//  * the unit of load-balanced work is an aligned 16-byte QUAD of col_indices: one segment search
//    and one 128-bit streaming load serve four neighbours, whose values are gathered with four
//    independent L1-allocating loads (hub values stay in L1, the index stream bypasses it) and
//    folded in registers -- the warp-wide segmented scan runs once per quad, not once per arc;
//  * a thread's VT quads are fetched first (all index loads in flight), then gathered (4*VT value
//    loads in flight), then reduced: three phases, no dependent load chain inside a phase;
//  * the 32 lanes of a warp hold 32 consecutive quads, so runs of equal segment are contiguous in
//    lane order; a ballot of the run heads gives every lane its distance to the head and the
//    scan needs only the 5 value shuffles.  The run tail writes: a row that lies entirely inside
//    the lane row is stored with a plain store (reduced[] was preset by the scan), anything else
//    (rows spanning lane rows, tiles, warps or CTAs) contributes one atomic per partial -- no
//    carry-out pass and no fix-up kernel;
//  * every warp walks its own pieces of the merged (quads U segment starts) list with a private
//    segment window in shared memory: no CTA barrier anywhere.
//
// reduced[] is pre-set by the quad scan (NeighborhoodQuads) to `identity` for empty
// neighbourhoods and to the operator's neutral element otherwise, so `identity` is never folded
// into a non-empty reduction (mgpu tests/test_segreduce.cu:40-58).  fp32 sum order is unspecified
// => tolerance-checked; min / max are exact.
#pragma once
#include "quad_advance.cuh"
#include "segreduce.cuh"

namespace b200 {

// Quad-count functor of the neighbourhood scan: leaves the row bounds beside the scan and presets reduced[].
template <class Value>
struct NeighborhoodQuads {
    const int *frontier;
    const uint32_t *offsets;
    uint2 *rows;
    Value *reduced;
    Value identity, neutral;
    int scatter;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const int v = frontier[i];
        uint32_t b = 0, e = 0;
        if (v >= 0) {
            b = __ldg(offsets + v);
            e = __ldg(offsets + v + 1);
        }
        rows[i] = make_uint2(b, e);
        if (v >= 0 || !scatter) reduced[scatter ? (uint32_t)v : i] = e > b ? neutral : identity;
        return e > b ? ((e - 1) >> 2) - (b >> 2) + 1u : 0u;
    }
};

#ifndef B200_QSEG_NT
#define B200_QSEG_NT 512
#endif
#ifndef B200_QSEG_VT
#define B200_QSEG_VT 2
#endif
#ifndef B200_QSEG_WSEG
#define B200_QSEG_WSEG 64
#endif
#ifndef B200_QSEG_MINB
#define B200_QSEG_MINB 2
#endif
#ifndef B200_QSEG_SPLIT
#define B200_QSEG_SPLIT 8
#endif

// ValueFn: Value operator()(int src, int nbr, uint32_t edge_id) -- plays Functor::get_value_to_reduce
// (pr_functor.hxx:27-29).  counters[B200_CNT_ARCS] += arcs reduced.
//
// HOT (b200_graph::hot_indices, see b200_frontier.h): the index array is the remapped copy in which an occurrence of
// hot vertex k is stored as ~k, and hot_vals[k] = the value of hot vertex k (evaluated once per launch by
// hot_values_kernel).  The CTA copies the table into shared memory; a negative index is served from there.  The
// plain kernel is bound by one L1 wavefront per divergent 4-byte gather (profiles/r01_ncu_segreduce_v2_quad.txt:
// l1tex 78 % busy, L1 hit rate 11 %); on RMAT graphs 40 K hot vertices take 40 % of the gathers off that path.
template <class Value, class ROp, class ValueFn, int NT, int VT, int WSEG, bool HOT = false, int MINB = B200_QSEG_MINB>
__global__ void __launch_bounds__(NT, MINB)
quad_segreduce_kernel(QuadArgs a, ValueFn vf, Value *__restrict__ reduced, int scatter, unsigned long long *counters,
                      const Value *__restrict__ hot_vals = nullptr, uint32_t hot_count = 0) {
    constexpr int NW = NT / 32;
    __shared__ uint32_t s_win[NW][4 * WSEG];
    __shared__ unsigned long long s_sum;
    extern __shared__ uint4 s_hot_raw[];
    Value *s_hot = reinterpret_cast<Value *>(s_hot_raw);
    if (HOT)
        for (uint32_t k = threadIdx.x; k < hot_count; k += NT) s_hot[k] = hot_vals[k];

    const unsigned lane = lane_id(), warp = threadIdx.x >> 5, le_mask = lanemask_lt() | (1u << lane);
    const unsigned long long Q = *a.total;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();

    uint32_t *start = s_win[warp];           // scanned quad start of staged segment j
    uint32_t *rb = start + WSEG;             // row begin (arc id)
    uint32_t *re = start + 2 * WSEG;         // row end
    int *vert = reinterpret_cast<int *>(start + 3 * WSEG);
    unsigned long long arc_cnt = 0;

    // one tile = up to 32*VT quads of the staged window, lane-strided: lane row i holds quads first_q + 32 i + lane
    auto tile = [&](uint32_t first_q, uint32_t nq, int j_lo, int j_hi, uint32_t first_seg) {
        int seg[VT], src[VT];
        uint32_t e0[VT], valid[VT];
        bool whole_head[VT], last[VT];
        int4 d[VT];
        // phase 1: locate + index loads
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t k = lane + 32u * i;
            seg[i] = -1;
            src[i] = 0;
            e0[i] = 0u;
            valid[i] = 0u;
            whole_head[i] = false;
            last[i] = false;
            d[i] = make_int4(0, 0, 0, 0);
            if (k < nq) {
                const uint32_t qa = first_q + k;
                const int j = lbs_locate(start, j_lo, j_hi, qa);
                const uint32_t b = rb[j], e = re[j];
                const uint32_t q = (b >> 2) + (qa - start[j]);
                d[i] = ld_stream_v4(a.indices4 + q);
                e0[i] = q << 2;
                const uint32_t lo = b > e0[i] ? b - e0[i] : 0u;
                const uint32_t hi = e - e0[i] < 4u ? e - e0[i] : 4u;
                valid[i] = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                seg[i] = j;
                src[i] = vert[j];
                whole_head[i] = qa == start[j];
                last[i] = q == ((e - 1u) >> 2);
            }
        }
        // phase 2: gather the neighbours' values (independent loads), fold the quad in registers
        Value x[VT];
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            auto value_of = [&](int nbr, uint32_t e) -> Value {
                if (HOT && nbr < 0) return s_hot[~nbr];
                return vf(src[i], nbr, e);
            };
            const Value v0 = (valid[i] & 1u) ? value_of(d[i].x, e0[i]) : ROp::neutral();
            const Value v1 = (valid[i] & 2u) ? value_of(d[i].y, e0[i] + 1u) : ROp::neutral();
            const Value v2 = (valid[i] & 4u) ? value_of(d[i].z, e0[i] + 2u) : ROp::neutral();
            const Value v3 = (valid[i] & 8u) ? value_of(d[i].w, e0[i] + 3u) : ROp::neutral();
            x[i] = ROp::apply(ROp::apply(v0, v1), ROp::apply(v2, v3));
            arc_cnt += __popc(valid[i]);
        }
        // phase 3: segmented scan over the 32 lanes of each lane row; the run tail writes
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const int j = seg[i];
            const int jp = __shfl_up_sync(FULL_MASK, j, 1);
            const unsigned heads = __ballot_sync(FULL_MASK, lane == 0 || jp != j);
            const unsigned whole = __ballot_sync(FULL_MASK, whole_head[i]);
            const unsigned hl = 31u - __clz(heads & le_mask);   // head lane of this lane's run
            const unsigned dist = lane - hl;
            Value v = x[i];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const Value vt = __shfl_up_sync(FULL_MASK, v, s);
                if (dist >= (unsigned)s) v = ROp::apply(vt, v);
            }
            const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
            if (j >= 0 && tail) {
                Value *dst = reduced + (scatter ? (uint32_t)src[i] : first_seg + (uint32_t)j);
                if (!scatter && last[i] && ((whole >> hl) & 1u)) *dst = v;   // the row's only partial
                else ROp::combine(dst, v);
            }
        }
    };

    // ---- this warp's pieces of the merged (quads U segment starts) list (as in quad_advance_kernel)
    const unsigned long long work = Q + a.num_segments;
    constexpr int SPLIT = B200_QSEG_SPLIT;
    static_assert(SPLIT >= 1 && SPLIT <= 16, "one lane per chunk boundary");
    const uint32_t total_warps = gridDim.x * NW, gw = warp * gridDim.x + blockIdx.x;
    unsigned long long chunk = ceil_div<unsigned long long>(work, (unsigned long long)total_warps * SPLIT);
    if (chunk < a.min_chunk) chunk = a.min_chunk;
    uint32_t sb = 0;
    {
        const unsigned long long piece = (unsigned long long)(lane >> 1) * total_warps + gw;
        unsigned long long dd = (piece + (lane & 1u)) * chunk;
        if (dd > work) dd = work;
        if (Q != 0 && lane < 2u * SPLIT) sb = merge_path_segments(a.scanned, a.num_segments, Q, dd);
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
                if (j_lo + 1 < ns && start[j_lo + 1] <= cur_q) j_lo = lbs_locate(start, j_lo + 1, ns, cur_q);
                int j_hi = j_lo + 1;
                if (j_hi < ns && start[j_hi] <= q_end - 1) j_hi = lbs_locate(start, j_hi, ns, q_end - 1) + 1;
                tile(cur_q, q_end - cur_q, j_lo, j_hi, cur_s);
                cur_q = q_end;
            }
            if (!more) break;                    // win_end == q1: piece done
            cur_s += (uint32_t)ns;               // scanned[cur_s + ns] == win_end == cur_q
            __syncwarp();
        }
        __syncwarp();
    }

#pragma unroll
    for (int s = 16; s > 0; s >>= 1) arc_cnt += __shfl_xor_sync(FULL_MASK, arc_cnt, s);
    if (lane == 0 && arc_cnt) atomicAdd(&s_sum, arc_cnt);
    __syncthreads();
    if (threadIdx.x == 0 && s_sum) atomicAdd(&counters[B200_CNT_ARCS], s_sum);
}

// hot_vals[k] = vf(-, hot_ids[k], -): the values the HOT kernel serves from shared memory
template <class Value, class ValueFn>
__global__ void hot_values_kernel(ValueFn vf, const int *__restrict__ hot_ids, uint32_t hot_count, Value *hot_vals) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < hot_count) hot_vals[k] = vf(hot_ids[k], hot_ids[k], 0u);
}

}  // namespace b200
