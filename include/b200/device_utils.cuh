// device_utils.cuh -- warp / memory helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() {
    unsigned r;
    asm("mov.u32 %0, %%laneid;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned r;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
    return r;
}

// Streaming read of data touched once per kernel (column indices, weights):
// read-only path, do not allocate in L1 so the probe working set keeps it.
__device__ __forceinline__ int ld_stream(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ld_stream_v2(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream_v4(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream_v4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Shared memory through explicit 32-bit shared-window addresses.  A generic pointer that has
// travelled through a struct makes the compiler rebuild the window base (S2R SR_CgaCtaId + LEA)
// in front of every access; these keep it a plain LDS/STS.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}

// visited |= "row has no in-arc" for the words first, first + stride, ... (ONE WORD PER THREAD: a word per warp
// iteration is a chain of dependent loads -- 0.27 ms for the 2 M words of a scale-26 bitmap).  iso: the prebuilt
// bitmap (b200_graph::no_in_arc_bitmap), else derived from the pull offsets.
// Nobody else may be updating `visited` meanwhile (the word is read through L2 and stored back).
__device__ __forceinline__ void or_no_in_arc_words(uint32_t first, uint32_t stride, const uint32_t *__restrict__ pull_offsets,
                                                   uint32_t n, const uint32_t *__restrict__ iso, uint32_t *visited) {
    const uint32_t num_words = (n + 31u) >> 5;
    if (iso) {   // four words per thread in flight (one at a time is a chain of dependent L2 round trips)
        for (uint32_t w0 = first; w0 < num_words; w0 += 4u * stride) {
            uint32_t m[4], v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t w = w0 + (uint32_t)k * stride;
                m[k] = w < num_words ? iso[w] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t w = w0 + (uint32_t)k * stride;
                v[k] = m[k] ? __ldcg(visited + w) : 0u;   // (through L2: a persistent kernel may hold a stale L1 line)
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (m[k]) visited[w0 + (uint32_t)k * stride] = v[k] | m[k];
        }
        return;
    }
    for (uint32_t w = first; w < num_words; w += stride) {
        uint32_t m = 0u;
        const uint32_t v0 = w << 5;
        uint32_t prev = pull_offsets[v0];
        for (uint32_t b = 0; b < 32u && v0 + b < n; ++b) {
            const uint32_t next = pull_offsets[v0 + b + 1];
            m |= (next == prev ? 1u : 0u) << b;
            prev = next;
        }
        if (m) visited[w] = __ldcg(visited + w) | m;
    }
}

template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace b200
