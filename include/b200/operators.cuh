// operators.cuh -- host-side launchers of the kernel templates.
//
// Instantiated (a) inside libb200_frontier.so with the built-in functors behind
// the C ABI and (b) in the caller's translation unit by include/gunrock/*.hxx
// with user functors (device functors cannot cross a C ABI).  Both use the
// ctx-owned b200_workspace, so no launcher allocates.
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include "advance.cuh"
#include "quad_advance.cuh"
#include "quad_segreduce.cuh"
#include "segreduce.cuh"
#include "tile_scan.cuh"
#include "workspace.h"

namespace b200 {

constexpr int SCAN_NT = 256, SCAN_VT = 4;
constexpr int COMPACT_NT = 256, COMPACT_VT = 8, BITLIST_VT = 4;
constexpr int LBS_NT = 256, LBS_VT = 8, LBS_SEG_T = 512;
constexpr uint32_t LBS_MIN_CHUNK = 2048;

inline cudaStream_t ws_stream(const b200_workspace *ws) { return (cudaStream_t)ws->stream; }

inline cudaError_t reset_counters(b200_workspace *ws) {
    return cudaMemsetAsync(ws->d_counters, 0, sizeof(unsigned long long) * B200_NUM_COUNTERS, ws_stream(ws));
}
// Device counters -> pinned host mirror; synchronises the ctx stream.
inline cudaError_t read_counters(b200_workspace *ws) {
    cudaError_t e = cudaMemcpyAsync(ws->h_counters, ws->d_counters, sizeof(unsigned long long) * B200_NUM_COUNTERS,
                                    cudaMemcpyDeviceToHost, ws_stream(ws));
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ws_stream(ws));
}

inline LookbackState next_lookback(b200_workspace *ws, uint32_t num_tiles) {
    ws->epoch = (ws->epoch + 1) & 0x3FFFFFFFu;
    if (ws->epoch == 0) {   // wrapped: stale entries could alias, wipe them
        cudaMemsetAsync(ws->d_status, 0, sizeof(unsigned long long) * (size_t)ws->status_tiles, ws_stream(ws));
        ws->epoch = 1;
    }
    LookbackState st;
    st.status = ws->d_status;
    st.tile_counter = ws->d_tile_counter;
    st.epoch = ws->epoch;
    st.num_tiles = num_tiles;
    return st;
}

// exclusive scan of fn(0..count) -> out, total -> *d_total.  count > 0.
template <class SizeFn>
cudaError_t launch_scan(b200_workspace *ws, SizeFn fn, uint32_t count, uint32_t *d_out, unsigned long long *d_total) {
    const uint32_t tiles = ceil_div<uint32_t>(count, SCAN_NT * SCAN_VT);
    if ((int64_t)tiles > ws->status_tiles) return cudaErrorInvalidValue;
    scan_sizes_kernel<SCAN_NT, SCAN_VT><<<tiles, SCAN_NT, 0, ws_stream(ws)>>>(fn, count, d_out, next_lookback(ws, tiles), d_total);
    ws->launches++;
    return cudaGetLastError();
}

template <class Pred>
cudaError_t launch_compact(b200_workspace *ws, Pred pred, uint32_t count, int *d_out, unsigned long long capacity,
                           unsigned long long *d_total, unsigned long long *d_overflow) {
    const uint32_t tiles = ceil_div<uint32_t>(count, COMPACT_NT * COMPACT_VT);
    if ((int64_t)tiles > ws->status_tiles) return cudaErrorInvalidValue;
    compact_kernel<COMPACT_NT, COMPACT_VT><<<tiles, COMPACT_NT, 0, ws_stream(ws)>>>(
        pred, count, d_out, capacity, next_lookback(ws, tiles), d_total, d_overflow);
    ws->launches++;
    return cudaGetLastError();
}

// bitmap (num_bits bits) -> ascending list of itf(idx) for every set bit idx.
template <class WordFn, class ItemFn>
cudaError_t launch_bitmap_list(b200_workspace *ws, WordFn wf, ItemFn itf, uint32_t num_bits, int *d_out, unsigned long long capacity,
                               unsigned long long *d_total, unsigned long long *d_overflow) {
    const uint32_t num_words = (num_bits + 31u) >> 5;
    const uint32_t tiles = ceil_div<uint32_t>(num_words, COMPACT_NT * BITLIST_VT);
    if ((int64_t)tiles > ws->status_tiles) return cudaErrorInvalidValue;
    bitmap_list_kernel<COMPACT_NT, BITLIST_VT><<<tiles, COMPACT_NT, 0, ws_stream(ws)>>>(
        wf, itf, num_words, d_out, capacity, next_lookback(ws, tiles), d_total, d_overflow, nullptr);
    ws->launches++;
    return cudaGetLastError();
}

// Persistent grid: resident CTAs per SM (queried once per instantiation) x SMs.
// Tag makes the cached occupancy unique per kernel instantiation (kernels that differ
// only in non-type template arguments share one function-pointer type).
template <class Tag, class Kernel>
int persistent_grid(Kernel k, int nt, const b200_workspace *ws) {
    static int per_sm = 0;   // per <Tag, Kernel> instantiation
    if (per_sm == 0) {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k, nt, 0) != cudaSuccess || b < 1) b = 1;
        per_sm = b;
    }
    return per_sm * ws->num_sms;
}

template <class Op, int OUT_MODE, bool DEG_SUM> struct AdvTag {};
template <class Value, class ROp, class ValueFn> struct SegTag {};

struct FrontierDegree {
    const int *frontier;
    const uint32_t *offsets;
    uint32_t row_shift;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const int v = frontier[i];
        if (v < 0) return 0u;
        const uint32_t r = (uint32_t)v >> row_shift;
        return __ldg(offsets + r + 1) - __ldg(offsets + r);
    }
};

inline LbsArgs make_lbs_args(const b200_workspace *ws, const int *d_frontier, uint32_t len,
                             const uint32_t *offsets, const int *indices, uint32_t row_shift = 0) {
    LbsArgs a;
    a.frontier = d_frontier;
    a.num_segments = len;
    a.scanned = ws->d_scanned;
    a.total = ws->d_counters + B200_CNT_TOTAL;
    a.offsets = offsets;
    a.indices = indices;
    a.min_chunk = LBS_MIN_CHUNK;
    a.row_shift = row_shift;
    return a;
}

// degree scan of the frontier into ws->d_scanned, m_F into counters[TOTAL].
inline cudaError_t launch_frontier_scan(b200_workspace *ws, const int *d_frontier, uint32_t len, const uint32_t *offsets,
                                        uint32_t row_shift = 0) {
    if ((int64_t)len > ws->scanned_capacity) return cudaErrorInvalidValue;
    FrontierDegree fn{d_frontier, offsets, row_shift};
    return launch_scan(ws, fn, len, ws->d_scanned, ws->d_counters + B200_CNT_TOTAL);
}

// LBS advance over a frontier whose degree scan is already in the workspace.
template <int OUT_MODE, bool DEG_SUM, class Op>
cudaError_t launch_lbs_advance(b200_workspace *ws, const LbsArgs &a, Op op, int *d_out, unsigned long long capacity,
                               const RoutedOut *routed = nullptr) {
    auto k = lbs_advance_kernel<Op, OUT_MODE, DEG_SUM, LBS_NT, LBS_VT, LBS_SEG_T>;
    const int grid = persistent_grid<AdvTag<Op, OUT_MODE, DEG_SUM>>(k, LBS_NT, ws);
    RoutedOut r;
    if (routed) r = *routed;
    else memset(&r, 0, sizeof r);
#ifdef B200_LBS_PAD_KB   // tuning builds: cap the resident CTAs per SM with unused dynamic shared memory
    k<<<ws->num_sms * (224 / B200_LBS_PAD_KB), LBS_NT, B200_LBS_PAD_KB * 1024 - 24 * 1024, ws_stream(ws)>>>(a, op, d_out, capacity, ws->d_counters, r);
#else
    k<<<grid, LBS_NT, 0, ws_stream(ws)>>>(a, op, d_out, capacity, ws->d_counters, r);
#endif
    ws->launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// quad advance (quad_advance.cuh): one CTA of QUAD_NT threads per SM; the per-warp windows,
// candidate buffers and output stages live in opt-in dynamic shared memory, the rest of the
// SM's SRAM stays L1 for the visited bitmap.
// (the B200_QUAD_* macros exist for tuning builds: python -m mini_b200.build --define ... --out ...)
// ---------------------------------------------------------------------------
#ifndef B200_QUAD_NT
#define B200_QUAD_NT 1024
#endif
#ifndef B200_QUAD_VT
#define B200_QUAD_VT 2
#endif
#ifndef B200_QUAD_WSEG
#define B200_QUAD_WSEG 64
#endif
constexpr int QUAD_NT = B200_QUAD_NT, QUAD_WSEG = B200_QUAD_WSEG, QUAD_WSTAGE = 64 * QUAD_R;
constexpr uint32_t QUAD_MIN_CHUNK = 64;            // work items per warp, lower bound

inline bool quad_aligned(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline QuadArgs make_quad_args(const b200_workspace *ws, const int *d_frontier, uint32_t len, const uint32_t *offsets,
                               const int *indices, const float *weights, uint32_t row_shift = 0) {
    QuadArgs a;
    a.frontier = d_frontier;
    a.num_segments = len;
    a.scanned = ws->d_scanned;
    a.rows = reinterpret_cast<const uint2 *>(ws->d_rows);
    a.total = ws->d_counters + B200_CNT_TOTAL;
    a.offsets = offsets;
    a.indices4 = reinterpret_cast<const int4 *>(indices);
    a.weights4 = reinterpret_cast<const float4 *>(weights);
    a.min_chunk = QUAD_MIN_CHUNK;
    a.row_shift = row_shift;
    a.dyn = nullptr;
    a.scanned_next = nullptr;
    a.rows_next = nullptr;
    return a;
}

// quad scan of the frontier into ws->d_scanned (+ row bounds into ws->d_rows), Q into counters[TOTAL].
inline cudaError_t launch_quad_scan(b200_workspace *ws, const int *d_frontier, uint32_t len, const uint32_t *offsets,
                                    uint32_t row_shift = 0) {
    if ((int64_t)len > ws->scanned_capacity) return cudaErrorInvalidValue;
    FrontierQuads fn{d_frontier, offsets, reinterpret_cast<uint2 *>(ws->d_rows), row_shift};
    return launch_scan(ws, fn, len, ws->d_scanned, ws->d_counters + B200_CNT_TOTAL);
}

// Quad advance over a frontier whose quad scan is already in the workspace.  m_F is left in
// counters[B200_CNT_ARCS], the emitted count in counters[B200_CNT_OUT] (or routed.count[]).
template <int OUT_MODE, bool DEG_SUM, class Op>
cudaError_t launch_quad_advance(b200_workspace *ws, const QuadArgs &a, Op op, int *d_out, unsigned long long capacity,
                                const RoutedOut *routed = nullptr) {
    constexpr int VT = Op::WEIGHTED ? 2 : B200_QUAD_VT;
    // segments staged per warp window: 64 for the unweighted operators (round-2 A/B, scale-22 push BFS: the 2 M-vertex
    // level 0.1522 -> 0.1446 ms, whole BFS -1.5 %; the heaviest level is unchanged), 32 for the weighted ones, whose
    // 3-word candidates already take 152 KB of shared memory (SSSP was 17 % slower with 64)
    constexpr int WSEG = Op::WEIGHTED ? (QUAD_WSEG < 32 ? QUAD_WSEG : 32) : QUAD_WSEG;
    constexpr bool STAGED = OUT_MODE == OUT_COMPACT || OUT_MODE == OUT_ROUTED;
    auto k = quad_advance_kernel<Op, OUT_MODE, DEG_SUM, QUAD_NT, VT, WSEG>;
    constexpr size_t smem = sizeof(uint32_t) * (QUAD_NT / 32) * quad_warp_words<Op, STAGED, VT, WSEG, QUAD_WSTAGE>();
    static_assert(smem <= 232448 - 64, "per-warp areas exceed the 227 KB opt-in shared memory of sm_100");
    static unsigned long long opted_in = 0;   // per <Op, OUT_MODE, DEG_SUM> instantiation, one bit per device
    if (!(opted_in >> (ws->device & 63) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
#ifdef B200_QUAD_CARVEOUT   // tuning builds: shared-memory share of the SM's SRAM in percent (default: the driver's choice)
        e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, B200_QUAD_CARVEOUT);
        if (e != cudaSuccess) return e;
#endif
        opted_in |= 1ull << (ws->device & 63);
    }
    RoutedOut r;
    if (routed) r = *routed;
    else memset(&r, 0, sizeof r);
    k<<<ws->num_sms * B200_QUAD_MINB, QUAD_NT, smem, ws_stream(ws)>>>(a, op, d_out, capacity, ws->d_counters, r);
    ws->launches++;
    return cudaGetLastError();
}

template <class Value, class ROp, class ValueFn>
cudaError_t launch_lbs_segreduce(b200_workspace *ws, const LbsArgs &a, ValueFn vf, Value *d_reduced, int scatter) {
    auto k = lbs_segreduce_kernel<Value, ROp, ValueFn, LBS_NT, LBS_VT, LBS_SEG_T>;
    const int grid = persistent_grid<SegTag<Value, ROp, ValueFn>>(k, LBS_NT, ws);
    k<<<grid, LBS_NT, 0, ws_stream(ws)>>>(a, vf, d_reduced, scatter);
    ws->launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// quad neighbourhood reduce (quad_segreduce.cuh)
// ---------------------------------------------------------------------------
constexpr int QSEG_NT = B200_QSEG_NT, QSEG_VT = B200_QSEG_VT, QSEG_WSEG = B200_QSEG_WSEG;

// quad scan of the frontier (+ row bounds into ws->d_rows, reduced[] preset), Q into counters[TOTAL].
template <class Value>
cudaError_t launch_neighborhood_quad_scan(b200_workspace *ws, const int *d_frontier, uint32_t len, const uint32_t *offsets,
                                          Value *d_reduced, Value identity, Value neutral, int scatter) {
    if ((int64_t)len > ws->scanned_capacity) return cudaErrorInvalidValue;
    NeighborhoodQuads<Value> fn{d_frontier, offsets, reinterpret_cast<uint2 *>(ws->d_rows), d_reduced, identity, neutral, scatter};
    return launch_scan(ws, fn, len, ws->d_scanned, ws->d_counters + B200_CNT_TOTAL);
}

// Reduce over a frontier whose neighbourhood quad scan is already in the workspace; the number of
// arcs reduced is left in counters[B200_CNT_ARCS].
template <class Value, class ROp, class ValueFn>
cudaError_t launch_quad_segreduce(b200_workspace *ws, const QuadArgs &a, ValueFn vf, Value *d_reduced, int scatter) {
    auto k = quad_segreduce_kernel<Value, ROp, ValueFn, QSEG_NT, QSEG_VT, QSEG_WSEG>;
    const int grid = persistent_grid<SegTag<Value, ROp, ValueFn>>(k, QSEG_NT, ws);
    k<<<grid, QSEG_NT, 0, ws_stream(ws)>>>(a, vf, d_reduced, scatter, ws->d_counters, (const Value *)nullptr, 0u);
    ws->launches++;
    return cudaGetLastError();
}

// The same over the remapped index copy of b200_graph::hot_indices (QuadArgs::indices4 must point at it): the values
// of the hot vertices are evaluated once into d_hot_vals and served from shared memory.  1 CTA x 1024 threads per SM.
constexpr int QSEG_HOT_MAX = 40960;   // == B200_HOT_MAX (b200_frontier.h): 160 KB of fp32 in shared memory
template <class Value, class ROp, class ValueFn>
cudaError_t launch_quad_segreduce_hot(b200_workspace *ws, const QuadArgs &a, ValueFn vf, Value *d_reduced, int scatter,
                                      const int *d_hot_ids, uint32_t hot_count, Value *d_hot_vals) {
    constexpr int NT = 1024;
    auto k = quad_segreduce_kernel<Value, ROp, ValueFn, NT, QSEG_VT, QSEG_WSEG, true, 1>;
    const size_t smem = sizeof(Value) * (size_t)hot_count;
    static unsigned long long opted_in = 0;   // per instantiation, one bit per device
    if (!(opted_in >> (ws->device & 63) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Value) * QSEG_HOT_MAX));
        if (e != cudaSuccess) return e;
        opted_in |= 1ull << (ws->device & 63);
    }
    hot_values_kernel<Value><<<(hot_count + 255) / 256, 256, 0, ws_stream(ws)>>>(vf, d_hot_ids, hot_count, d_hot_vals);
    k<<<ws->num_sms, NT, smem, ws_stream(ws)>>>(a, vf, d_reduced, scatter, ws->d_counters, d_hot_vals, hot_count);
    ws->launches += 2;
    return cudaGetLastError();
}

}  // namespace b200
