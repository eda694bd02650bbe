// tile_scan.cuh -- single-pass (decoupled look-back) scan and stable compaction.
//
// Replaces mgpu transform_scan (kernel_scan.hxx:14-157: upsweep + spine +
// downsweep, the input lambda evaluated twice) and transform_compact
// (kernel_compact.hxx:9-137: upsweep + scan + pinned-host read + downsweep) with
// ONE kernel each: every tile publishes its aggregate, looks back over its
// predecessors' status words and writes its results in the same pass, so the
// input functor runs exactly once per item and nothing is re-read.
//
// Items are held warp-striped (tile item = warp*32*VT + i*32 + lane) so global
// reads and writes of consecutive lanes are consecutive addresses.
#pragma once
#include "device_utils.cuh"
#include "loop_dyn.cuh"

namespace b200 {

struct LookbackState {
    unsigned long long *status;   // [num_tiles] (flag:2 | epoch:30 | value:32)
    unsigned int *tile_counter;   // dynamic tile ids => predecessors are always resident
    unsigned int epoch;           // entries with another epoch are "not ready"
    unsigned int num_tiles;
};

constexpr unsigned long long LB_FLAG_AGG = 1ull, LB_FLAG_PREFIX = 2ull;
__device__ __forceinline__ unsigned long long lb_pack(unsigned long long flag, unsigned epoch, uint32_t value) {
    return (flag << 62) | ((unsigned long long)(epoch & 0x3FFFFFFFu) << 32) | value;
}

// CTA-wide exclusive scan of VT values per thread in warp-striped order.
template <int NT, int VT>
struct TileScan {
    static constexpr int NW = NT / 32;
    static constexpr int NV = NT * VT;
    struct Smem { uint32_t warp_total[NW]; };

    // returns the tile total; ex[i] = sum of all tile items that precede item i of this thread
    __device__ static __forceinline__ uint32_t run(const uint32_t (&v)[VT], uint32_t (&ex)[VT], Smem &s) {
        const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
        uint32_t running = 0;
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            uint32_t inc = v[i];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(FULL_MASK, inc, d);
                if (lane >= (unsigned)d) inc += t;
            }
            ex[i] = running + inc - v[i];
            running += __shfl_sync(FULL_MASK, inc, 31);
        }
        if (lane == 0) s.warp_total[warp] = running;
        __syncthreads();
        uint32_t wprefix = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t t = s.warp_total[w];
            if ((unsigned)w < warp) wprefix += t;
            total += t;
        }
#pragma unroll
        for (int i = 0; i < VT; ++i) ex[i] += wprefix;
        return total;
    }
};

__device__ __forceinline__ uint32_t claim_tile(const LookbackState &st, uint32_t *s_tile) {
    if (threadIdx.x == 0) {
        uint32_t t = atomicAdd(st.tile_counter, 1u);
        if (t == st.num_tiles - 1) *st.tile_counter = 0;   // grid == num_tiles: nobody claims after this
        *s_tile = t;
    }
    __syncthreads();
    return *s_tile;
}

// Publishes tile_total, returns the sum of the totals of all tiles < tile.  Whole CTA must call.
__device__ __forceinline__ uint32_t lookback_exclusive(const LookbackState &st, uint32_t tile,
                                                       uint32_t tile_total, uint32_t *s_bcast) {
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        if (tile == 0) {
            if (lane == 0) {
                st_volatile_u64(st.status, lb_pack(LB_FLAG_PREFIX, st.epoch, tile_total));
                *s_bcast = 0;
            }
        } else {
            if (lane == 0) st_volatile_u64(st.status + tile, lb_pack(LB_FLAG_AGG, st.epoch, tile_total));
            uint32_t excl = 0;
            int pred = (int)tile - 1;
            const unsigned ep = st.epoch & 0x3FFFFFFFu;
            for (;;) {
                const int idx = pred - (int)lane;
                unsigned long long w;
                bool ready;
                do {
                    w = idx >= 0 ? ld_volatile_u64(st.status + idx) : lb_pack(LB_FLAG_PREFIX, ep, 0);
                    ready = ((unsigned)(w >> 32) & 0x3FFFFFFFu) == ep && (w >> 62) != 0;
                } while (!__all_sync(FULL_MASK, ready));
                const unsigned pm = __ballot_sync(FULL_MASK, (w >> 62) == LB_FLAG_PREFIX);
                const uint32_t val = (uint32_t)w;
                if (pm) {
                    const unsigned first = __ffs(pm) - 1;   // nearest predecessor holding an inclusive prefix
                    excl += warp_sum(lane <= first ? val : 0u);
                    break;
                }
                excl += warp_sum(val);
                pred -= 32;
            }
            if (lane == 0) {
                st_volatile_u64(st.status + tile, lb_pack(LB_FLAG_PREFIX, st.epoch, excl + tile_total));
                *s_bcast = excl;
            }
        }
    }
    __syncthreads();
    return *s_bcast;
}

// out[i] = sum_{k<i} fn(k) for i in [0,count); *total_out = sum of all.  fn runs once per item.
template <int NT, int VT, class SizeFn>
__global__ void __launch_bounds__(NT) scan_sizes_kernel(SizeFn fn, uint32_t count, uint32_t *__restrict__ out,
                                                        LookbackState st, unsigned long long *total_out) {
    using TS = TileScan<NT, VT>;
    __shared__ typename TS::Smem sm;
    __shared__ uint32_t s_tile, s_bcast;
    const uint32_t tile = claim_tile(st, &s_tile);
    const uint32_t base = tile * TS::NV + (threadIdx.x >> 5) * (32 * VT) + lane_id();
    uint32_t v[VT], ex[VT];
#pragma unroll
    for (int i = 0; i < VT; ++i) {
        const uint32_t idx = base + i * 32;
        v[i] = idx < count ? fn(idx) : 0u;
    }
    const uint32_t total = TS::run(v, ex, sm);
    const uint32_t excl = lookback_exclusive(st, tile, total, &s_bcast);
#pragma unroll
    for (int i = 0; i < VT; ++i) {
        const uint32_t idx = base + i * 32;
        if (idx < count) out[idx] = excl + ex[i];
    }
    if (tile == st.num_tiles - 1 && threadIdx.x == 0) *total_out = (unsigned long long)excl + total;
}

// Stable compaction: for idx in [0,count): if pred(idx, item) then out[rank] = item.
// pred runs exactly once per idx (it may mutate problem data like the reference's
// cond_filter does).  Writes past `capacity` are dropped and flagged.
template <int NT, int VT, class Pred>
__global__ void __launch_bounds__(NT) compact_kernel(Pred pred, uint32_t count, int *__restrict__ out,
                                                     unsigned long long capacity, LookbackState st,
                                                     unsigned long long *total_out,
                                                     unsigned long long *overflow_flag) {
    using TS = TileScan<NT, VT>;
    __shared__ typename TS::Smem sm;
    __shared__ uint32_t s_tile, s_bcast;
    const uint32_t tile = claim_tile(st, &s_tile);
    const uint32_t base = tile * TS::NV + (threadIdx.x >> 5) * (32 * VT) + lane_id();
    uint32_t v[VT], ex[VT];
    int item[VT];
#pragma unroll
    for (int i = 0; i < VT; ++i) {
        const uint32_t idx = base + i * 32;
        item[i] = -1;
        v[i] = (idx < count && pred(idx, item[i])) ? 1u : 0u;
    }
    const uint32_t total = TS::run(v, ex, sm);
    const uint32_t excl = lookback_exclusive(st, tile, total, &s_bcast);
    bool over = false;
#pragma unroll
    for (int i = 0; i < VT; ++i) {
        if (v[i]) {
            const unsigned long long dest = (unsigned long long)excl + ex[i];
            if (dest < capacity) out[dest] = item[i];
            else over = true;
        }
    }
    if (over) *overflow_flag = 1ull;
    if (tile == st.num_tiles - 1 && threadIdx.x == 0) *total_out = (unsigned long long)excl + total;
}

// ---------------------------------------------------------------------------
// Graph-driven level loop forms (loop_dyn.cuh): the item count and the look-back tag are read
// from device memory, so the grid cannot be sized by the count -- a fixed persistent grid claims
// tiles from the dynamic counter until they run out (predecessors of a claimed tile are always
// resident, so the look-back cannot deadlock).  The counter is zeroed by the level's decide kernel.
// ---------------------------------------------------------------------------
template <int NT, int VT, class SizeFn>
__global__ void __launch_bounds__(NT) scan_sizes_dyn_kernel(SizeFn fn, const LoopDyn *dyn, uint32_t run_bit,
                                                            uint32_t *__restrict__ out, unsigned long long *status,
                                                            unsigned int *tile_counter, unsigned long long *total_out) {
    using TS = TileScan<NT, VT>;
    __shared__ typename TS::Smem sm;
    __shared__ uint32_t s_tile, s_bcast;
    loop_trace(dyn, 2);
    if (!(dyn->run & run_bit)) return;
    if (dyn->scanned_in) out = dyn->scanned_in;   // work-creating loop: this level's scan buffer is chosen on the device
    const uint32_t count = dyn->len;
    LookbackState st;
    st.status = status;
    st.tile_counter = tile_counter;
    st.epoch = dyn->epoch;
    st.num_tiles = (count + TS::NV - 1) / TS::NV;
    if (count == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *total_out = 0ull;
        return;
    }
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= st.num_tiles) break;
        const uint32_t base = tile * TS::NV + (threadIdx.x >> 5) * (32 * VT) + lane_id();
        uint32_t v[VT], ex[VT];
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t idx = base + i * 32;
            v[i] = idx < count ? fn(idx) : 0u;
        }
        const uint32_t total = TS::run(v, ex, sm);
        const uint32_t excl = lookback_exclusive(st, tile, total, &s_bcast);
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t idx = base + i * 32;
            if (idx < count) out[idx] = excl + ex[i];
        }
        if (tile == st.num_tiles - 1 && threadIdx.x == 0) *total_out = (unsigned long long)excl + total;
        __syncthreads();   // s_tile, s_bcast and the scan scratch are reused by the next tile
    }
}

// Stable compaction of idx in [0,count) into the list dyn->in (the frontier the next push level reads).
// clear_words (nullable): a bitmap of count bits that is zeroed word by word once the word's 32 indices
// have been tested (a warp's lanes hold exactly one word per item row).
template <int NT, int VT, class Pred>
__global__ void __launch_bounds__(NT) compact_dyn_kernel(Pred pred, uint32_t count, const LoopDyn *dyn, uint32_t run_bit,
                                                         unsigned long long capacity, unsigned long long *status,
                                                         unsigned int *tile_counter, unsigned long long *total_out,
                                                         unsigned long long *overflow_flag, uint32_t *clear_words) {
    using TS = TileScan<NT, VT>;
    __shared__ typename TS::Smem sm;
    __shared__ uint32_t s_tile, s_bcast;
    if (!(dyn->run & run_bit)) return;
    int *out = const_cast<int *>(dyn->in);
    LookbackState st;
    st.status = status;
    st.tile_counter = tile_counter;
    st.epoch = (dyn->epoch + 1u) & 0x3FFFFFFFu;
    st.num_tiles = (count + TS::NV - 1) / TS::NV;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= st.num_tiles) break;
        const uint32_t base = tile * TS::NV + (threadIdx.x >> 5) * (32 * VT) + lane_id();
        uint32_t v[VT], ex[VT];
        int item[VT];
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t idx = base + i * 32;
            item[i] = -1;
            v[i] = (idx < count && pred(idx, item[i])) ? 1u : 0u;
        }
        if (clear_words) {
            __syncwarp();
            if (lane_id() == 0) {
#pragma unroll
                for (int i = 0; i < VT; ++i)
                    if (base + i * 32 < count) clear_words[(base + i * 32) >> 5] = 0u;
            }
        }
        const uint32_t total = TS::run(v, ex, sm);
        const uint32_t excl = lookback_exclusive(st, tile, total, &s_bcast);
        bool over = false;
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            if (v[i]) {
                const unsigned long long dest = (unsigned long long)excl + ex[i];
                if (dest < capacity) out[dest] = item[i];
                else over = true;
            }
        }
        if (over) *overflow_flag = 1ull;
        if (tile == st.num_tiles - 1 && threadIdx.x == 0) *total_out = (unsigned long long)excl + total;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Bitmap -> sorted list, one 32-bit WORD per thread item (popc + scan + bit walk) instead of one bit
// per item through compact_kernel: 32x fewer scan items.  This is the pull -> push hand-over of the
// direction-optimising BFS (the reference's dense frontier is an int per vertex, bfs_enactor.hxx:83,
// and it never converts back); at scale 26 the bit-wise form took 0.24 ms of a 1.1 ms traversal.
//   WordFn: uint32_t operator()(uint32_t w)   the w-th word of the bitmap
//   ItemFn: int operator()(uint32_t idx)      the vertex id of bit idx
// Stable (ascending idx).  clear_words (nullable): zeroed behind the read.
// ---------------------------------------------------------------------------
struct BitmapWords {          // a plain bitmap
    const uint32_t *bm;
    __device__ __forceinline__ uint32_t operator()(uint32_t w) const { return bm[w]; }
};
struct BitmapWordsDyn {       // one of two bitmaps, selected on the device (LoopDyn::bsel)
    const LoopDyn *dyn;
    const uint32_t *bm0, *bm1;
    __device__ __forceinline__ uint32_t operator()(uint32_t w) const { return (dyn->bsel ? bm1 : bm0)[w]; }
};
struct IdentityItem {
    __device__ __forceinline__ int operator()(uint32_t idx) const { return (int)idx; }
};

template <int NT, int VT, class WordFn, class ItemFn>
__device__ __forceinline__ void bitmap_list_tile(const WordFn &wf, const ItemFn &itf, uint32_t num_words, uint32_t tile,
                                                 const LookbackState &st, int *__restrict__ out, unsigned long long capacity,
                                                 unsigned long long *total_out, unsigned long long *overflow_flag,
                                                 uint32_t *clear_words, typename TileScan<NT, VT>::Smem &sm, uint32_t *s_bcast) {
    using TS = TileScan<NT, VT>;
    const uint32_t base = tile * TS::NV + (threadIdx.x >> 5) * (32 * VT) + lane_id();
    uint32_t word[VT], c[VT], ex[VT];
#pragma unroll
    for (int i = 0; i < VT; ++i) {
        const uint32_t w = base + i * 32;
        word[i] = w < num_words ? wf(w) : 0u;
        c[i] = __popc(word[i]);
        if (clear_words && w < num_words) clear_words[w] = 0u;
    }
    const uint32_t total = TS::run(c, ex, sm);
    const uint32_t excl = lookback_exclusive(st, tile, total, s_bcast);
    bool over = false;
#pragma unroll
    for (int i = 0; i < VT; ++i) {
        uint32_t bits = word[i];
        unsigned long long dest = (unsigned long long)excl + ex[i];
        const uint32_t idx0 = (base + i * 32) << 5;
        while (bits) {
            const uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (dest < capacity) out[dest] = itf(idx0 + b);
            else over = true;
            ++dest;
        }
    }
    if (over) *overflow_flag = 1ull;
    if (tile == st.num_tiles - 1 && threadIdx.x == 0) *total_out = (unsigned long long)excl + total;
}

template <int NT, int VT, class WordFn, class ItemFn>
__global__ void __launch_bounds__(NT) bitmap_list_kernel(WordFn wf, ItemFn itf, uint32_t num_words, int *__restrict__ out,
                                                         unsigned long long capacity, LookbackState st,
                                                         unsigned long long *total_out, unsigned long long *overflow_flag,
                                                         uint32_t *clear_words) {
    __shared__ typename TileScan<NT, VT>::Smem sm;
    __shared__ uint32_t s_tile, s_bcast;
    const uint32_t tile = claim_tile(st, &s_tile);
    bitmap_list_tile<NT, VT>(wf, itf, num_words, tile, st, out, capacity, total_out, overflow_flag, clear_words, sm, &s_bcast);
}

// graph-driven loop form: writes the list dyn->in; persistent grid, look-back tag dyn->epoch + 1
template <int NT, int VT, class WordFn, class ItemFn>
__global__ void __launch_bounds__(NT) bitmap_list_dyn_kernel(WordFn wf, ItemFn itf, uint32_t num_words, const LoopDyn *dyn,
                                                             uint32_t run_bit, unsigned long long capacity,
                                                             unsigned long long *status, unsigned int *tile_counter,
                                                             unsigned long long *total_out, unsigned long long *overflow_flag,
                                                             uint32_t *clear_words) {
    using TS = TileScan<NT, VT>;
    __shared__ typename TS::Smem sm;
    __shared__ uint32_t s_tile, s_bcast;
    loop_trace(dyn, 11);
    if (!(dyn->run & run_bit)) return;
    int *out = const_cast<int *>(dyn->in);
    LookbackState st;
    st.status = status;
    st.tile_counter = tile_counter;
    st.epoch = (dyn->epoch + 1u) & 0x3FFFFFFFu;
    st.num_tiles = (num_words + TS::NV - 1) / TS::NV;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= st.num_tiles) break;
        bitmap_list_tile<NT, VT>(wf, itf, num_words, tile, st, out, capacity, total_out, overflow_flag, clear_words, sm, &s_bcast);
        __syncthreads();
    }
}

}  // namespace b200
