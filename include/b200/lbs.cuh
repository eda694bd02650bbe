// lbs.cuh -- merge-path load-balanced search over (arcs x frontier segments).
//
// Replaces mgpu load_balance_partitions + transform_lbs
// (kernel_load_balance.hxx:10-62, cta_load_balance.hxx:17-183, search.hxx:14-37):
//   * the work list is the merge of the arc indices 0..m_F-1 with the scanned
//     segment starts; it is cut into gridDim.x equal CONTIGUOUS chunks, so a CTA
//     does one merge-path binary search for each end of its chunk (no separate
//     partition kernel, no `mp` allocation) and then walks its chunk tile by tile;
//   * the segment starts, the source vertex and (row_offset - start) of up to SEG_T
//     segments are staged in shared memory once per window; an arc finds its segment
//     with a binary search over only the staged segments its tile touches;
//   * arcs are assigned to threads strided by NT, so the 32 lanes of a warp read
//     32 consecutive column indices (one 128-byte line) whenever they sit in the
//     same segment;
//   * m_F is read from device memory (written by the degree scan), so the host
//     never has to know it to size the launch: the grid is persistent.
#pragma once
#include "device_utils.cuh"

namespace b200 {

struct LbsArgs {
    const int *frontier;               // [num_segments] vertex ids
    uint32_t num_segments;             // |F|
    const uint32_t *scanned;           // [num_segments] exclusive scan of degrees
    const unsigned long long *total;   // device: m_F
    const uint32_t *offsets;           // row (push) or column (pull) offsets [n+1]
    const int *indices;                // col (push) or row (pull) indices [m]
    uint32_t min_chunk;                // lower bound on the work items per CTA
    uint32_t row_shift;                // row of vertex v in `offsets` is v >> row_shift (cyclic 1D partition: log2 P)
};

template <int NT, int VT, int SEG_T>
struct LbsSmem {
    uint32_t start[SEG_T];   // scanned start of staged segment j
    uint32_t base[SEG_T];    // offsets[vertex_j] - start_j  (so edge_id = base + arc)
    int vert[SEG_T];         // vertex_j
    uint32_t bounds[4];      // s0, a0, s1, a1
};

// Number of segment starts that precede diagonal d in the merged order
// ("segment start s goes before arc a iff scanned[s] <= a").
__device__ __forceinline__ uint32_t merge_path_segments(const uint32_t *scanned, uint32_t num_segments,
                                                        unsigned long long m, unsigned long long d) {
    unsigned long long lo = d > m ? d - m : 0ull;
    unsigned long long hi = d < (unsigned long long)num_segments ? d : (unsigned long long)num_segments;
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if ((unsigned long long)__ldg(scanned + mid) <= d - 1 - mid) lo = mid + 1;
        else hi = mid;
    }
    return (uint32_t)lo;
}

// Largest j in [lo, hi) with start[j] <= arc.  Requires start[lo] <= arc.
__device__ __forceinline__ int lbs_locate(const uint32_t *start, int lo, int hi, uint32_t arc) {
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (start[mid] <= arc) lo = mid;
        else hi = mid;
    }
    return lo;
}

// Walks this CTA's chunk.  Segments are staged one WINDOW (up to SEG_T segments) at a
// time; a window is cut into tiles of up to NT*VT arcs.  For every tile calls
//     body(first_arc, n_arcs, j_lo, j_hi, first_seg)
// where the staged segments [j_lo, j_hi) (window-relative; slot = first_seg + j) cover
// the tile's arcs and sm.start[j_lo] <= first_arc.  There is NO CTA barrier between the
// tiles of a window (the staged slice is read-only, every thread derives the tile
// bounds itself), so warps run ahead of each other freely; the only barriers are the
// two around a window refill.  body must not use CTA-wide barriers.
template <int NT, int VT, int SEG_T, class Body>
__device__ __forceinline__ void lbs_for_each_tile(const LbsArgs &a, LbsSmem<NT, VT, SEG_T> &sm, Body body) {
    constexpr uint32_t ARC_T = NT * VT;
    const unsigned long long m = *a.total;
    const unsigned long long work = m + a.num_segments;
    unsigned long long chunk = ceil_div<unsigned long long>(work, gridDim.x);
    if (chunk < a.min_chunk) chunk = a.min_chunk;
    const unsigned long long d0 = (unsigned long long)blockIdx.x * chunk;
    if (d0 >= work || m == 0) return;
    const unsigned long long d1 = d0 + chunk < work ? d0 + chunk : work;
    if (threadIdx.x == 0 || threadIdx.x == 32) {
        const unsigned long long d = threadIdx.x == 0 ? d0 : d1;
        const uint32_t s = merge_path_segments(a.scanned, a.num_segments, m, d);
        sm.bounds[threadIdx.x ? 2 : 0] = s;
        sm.bounds[threadIdx.x ? 3 : 1] = (uint32_t)(d - s);
    }
    __syncthreads();
    const uint32_t s0 = sm.bounds[0], a0 = sm.bounds[1], s1 = sm.bounds[2], a1 = sm.bounds[3];
    if (a0 >= a1) return;
    uint32_t cur_s = s0 > 0 ? s0 - 1 : 0;   // segment that contains arc a0
    uint32_t cur_a = a0;
    while (cur_a < a1) {                     // one iteration per window
        const int ns = (int)min((uint32_t)SEG_T, s1 - cur_s);
        for (int j = threadIdx.x; j < ns; j += NT) {
            const uint32_t st = __ldg(a.scanned + cur_s + j);
            const int v = __ldg(a.frontier + cur_s + j);
            sm.start[j] = st;
            sm.vert[j] = v;
            sm.base[j] = (v >= 0 ? __ldg(a.offsets + (v >> a.row_shift)) : 0u) - st;
        }
        __syncthreads();
        // arcs of segments that are not staged wait for the next window
        uint32_t win_end = a1;
        if (cur_s + (uint32_t)ns < s1) {
            const uint32_t lim = __ldg(a.scanned + cur_s + ns);
            if (lim < win_end) win_end = lim;
        }
        int j_lo = 0;                        // sm.start[0] <= cur_a by construction
        while (cur_a < win_end) {
            const uint32_t a_end = win_end - cur_a > ARC_T ? cur_a + ARC_T : win_end;
            j_lo = lbs_locate(sm.start, j_lo, ns, cur_a);
            const int j_hi = lbs_locate(sm.start, j_lo, ns, a_end - 1) + 1;
            body(cur_a, a_end - cur_a, j_lo, j_hi, cur_s);
            cur_a = a_end;
        }
        // the next window starts at the segment that contains arc win_end
        cur_s += (uint32_t)lbs_locate(sm.start, j_lo, ns, win_end);
        __syncthreads();                     // everyone is done reading this window
    }
}

}  // namespace b200
