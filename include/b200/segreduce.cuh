// segreduce.cuh -- warp-segmented neighborhood reduce over the LBS walk.
//
// Replaces mgpu lbs_segreduce (kernel_segreduce.hxx:306-387: thread-serial reduce +
// CTA segscan + recursive fix-up kernel) as used by neighborhood_kernel
// (neighborhood.hxx:47-58).  The 32 lanes of a warp hold 32 consecutive arcs, so
// runs of equal segment are contiguous in lane order: a 5-step shuffle segmented
// scan leaves each run's value in its last lane, which combines it into
// reduced[slot] with one atomic (a segment spanning several warp rows, tiles or
// CTAs simply contributes several partials; no carry-out pass and no fix-up
// kernel).  reduced[] is pre-set by the degree scan to `identity` for empty
// neighbourhoods and to the operator's neutral element otherwise, so `identity`
// is never folded into a non-empty reduction (mgpu tests/test_segreduce.cu:40-58).
// Sum order is unspecified => fp32 results are tolerance-checked.
#pragma once
#include <cfloat>
#include "device_utils.cuh"
#include "lbs.cuh"

namespace b200 {

struct PlusF32 {
    __host__ __device__ static float neutral() { return 0.0f; }
    __device__ static float apply(float a, float b) { return a + b; }
    __device__ static void combine(float *p, float v) { atomicAdd(p, v); }
};
struct MinF32 {
    __host__ __device__ static float neutral() { return FLT_MAX; }
    __device__ static float apply(float a, float b) { return fminf(a, b); }
    __device__ static void combine(float *p, float v) {
        int *ip = reinterpret_cast<int *>(p);
        int old = *ip;
        while (v < __int_as_float(old)) {
            const int assumed = old;
            old = atomicCAS(ip, assumed, __float_as_int(v));
            if (old == assumed) break;
        }
    }
};
struct MaxF32 {
    __host__ __device__ static float neutral() { return -FLT_MAX; }
    __device__ static float apply(float a, float b) { return fmaxf(a, b); }
    __device__ static void combine(float *p, float v) {
        int *ip = reinterpret_cast<int *>(p);
        int old = *ip;
        while (v > __int_as_float(old)) {
            const int assumed = old;
            old = atomicCAS(ip, assumed, __float_as_int(v));
            if (old == assumed) break;
        }
    }
};
struct PlusI32 {
    __host__ __device__ static int neutral() { return 0; }
    __device__ static int apply(int a, int b) { return a + b; }
    __device__ static void combine(int *p, int v) { atomicAdd(p, v); }
};
struct MinI32 {
    __host__ __device__ static int neutral() { return INT_MAX; }
    __device__ static int apply(int a, int b) { return a < b ? a : b; }
    __device__ static void combine(int *p, int v) { atomicMin(p, v); }
};
struct MaxI32 {
    __host__ __device__ static int neutral() { return INT_MIN; }
    __device__ static int apply(int a, int b) { return a > b ? a : b; }
    __device__ static void combine(int *p, int v) { atomicMax(p, v); }
};

// Degree functor of the neighbourhood scan: also presets reduced[].
template <class Value>
struct NeighborhoodDegree {
    const int *frontier;
    const uint32_t *offsets;
    Value *reduced;
    Value identity, neutral;
    int scatter;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const int v = frontier[i];
        const uint32_t d = v >= 0 ? offsets[v + 1] - offsets[v] : 0u;
        if (v >= 0 || !scatter) reduced[scatter ? (uint32_t)v : i] = d ? neutral : identity;
        return d;
    }
};

// ValueFn: Value operator()(int src, int nbr, uint32_t edge_id) -- plays
// Functor::get_value_to_reduce (pr_functor.hxx:27-29).
template <class Value, class ROp, class ValueFn, int NT, int VT, int SEG_T>
__global__ void __launch_bounds__(NT) lbs_segreduce_kernel(LbsArgs a, ValueFn vf, Value *__restrict__ reduced,
                                                           int scatter) {
    __shared__ LbsSmem<NT, VT, SEG_T> sm;
    lbs_for_each_tile<NT, VT, SEG_T>(a, sm, [&](uint32_t first_arc, uint32_t n_arcs, int j_lo, int j_hi, uint32_t first_seg) {
        const unsigned lane = lane_id();
        int seg[VT];
        Value x[VT];
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t k = threadIdx.x + i * NT;
            seg[i] = -1;
            x[i] = ROp::neutral();
            if (k < n_arcs) {
                const uint32_t arc = first_arc + k;
                const int j = lbs_locate(sm.start, j_lo, j_hi, arc);
                const uint32_t eid = sm.base[j] + arc;
                seg[i] = j;
                x[i] = vf(sm.vert[j], ld_stream(a.indices + eid), eid);
            }
        }
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            Value v = x[i];
            const int j = seg[i];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Value vt = __shfl_up_sync(FULL_MASK, v, d);
                const int jt = __shfl_up_sync(FULL_MASK, j, d);
                if (lane >= (unsigned)d && jt == j) v = ROp::apply(vt, v);
            }
            const int jn = __shfl_down_sync(FULL_MASK, j, 1);
            if (j >= 0 && (lane == 31 || jn != j))
                ROp::combine(reduced + (scatter ? (uint32_t)sm.vert[j] : first_seg + (uint32_t)j), v);
        }
    });
}

}  // namespace b200
