// near_far.cuh -- the two passes that open the next bucket of the near-far SSSP loop (NearFar, loop_dyn.cuh).
//
// sssp_enactor_t::enact (sssp_enactor.hxx:40-72) expands every improved vertex in the next iteration; on RMAT-22 with
// weights 1..64 that relaxes 242 M arcs for 134 M reached.  b200_sssp_run keeps the reference's
// operators (advance with the relax functor, fused filter) and changes only WHICH improved vertices form the next
// frontier: those below `cutoff`.  When that near frontier runs dry every distance below `cutoff` is final, and
//
//   sssp_pending_min_kernel   lo = the smallest finite distance >= cutoff (none: the traversal is done); the last CTA
//                             to finish widens / resets the bucket width and publishes [lo, lo + delta)
//   sssp_take_kernel          every vertex with lo <= dist < lo + delta becomes the near frontier (none of them can
//                             have been expanded: expansions so far happened below the old cutoff and distances only
//                             fall); the last CTA publishes the count
//
// There is no far pile: a pending vertex is recorded by its distance alone, so the advance kernel writes nothing for
// it and stale pile entries cannot exist.  Both passes stream the n distances (scale 22: 16 MB, L2-resident).
//
// Bucket width: delta0 = 3.5 * mean weight / mean degree (a strided sample of the weights, once per graph).  A bucket
// that expanded fewer than m/32 arcs doubles the width (x4 once a bucket above m/4 has been seen); a bucket above m/4
// resets it to delta0; and once 60 % of the arcs have been expanded the rest is ONE bucket: the tail of the distance
// distribution is sparse, every bucket change costs two passes plus an iteration of ~17 us whatever it holds, and the
// reference's order on a remainder of low-degree vertices re-expands little.  Measured policy, not theory (B200,
// RMAT-22, integer weights 1..64; profiles/r02_sssp_near_far.txt): the reference's order 8-9 iterations, 1.80x the
// reached arcs, 1.57 ms; buckets all the way 19 iterations in 6 buckets, 1.006x, 1.28 ms; with the last-bucket rule
// 9 iterations in 2 buckets, 1.019x, 1.01 ms.  Any width gives the same distances.  Weights must be non-negative (as
// the integer atomicMin of the relax already requires).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b200/loop_dyn.cuh"

namespace b200 {

constexpr int NF_NT = 256;
constexpr unsigned NF_INF_BITS = 0x7f7fffffu;   // FLT_MAX: the reference's "unreached" (sssp_problem.hxx:44)

inline unsigned nf_grid(int num_sms, int64_t n) {   // both passes: one float4 per thread per round
    const int64_t rounds = (((n + 3) >> 2) + NF_NT - 1) / NF_NT;
    const int64_t full = (int64_t)num_sms * 8;
    return (unsigned)(rounds < 1 ? 1 : rounds < full ? rounds : full);
}

static __global__ void nf_init_kernel(NearFar *nf, float delta0, long long m) { near_far_reset(nf, delta0, m); }

// One thread, after the last CTA of the min pass: choose the next bucket.
__device__ __forceinline__ bool nf_open_bucket(NearFar *nf) {
    const unsigned bits = *reinterpret_cast<volatile unsigned *>(&nf->pending_min);
    nf->pending_min = NF_INF_BITS;
    nf->ticket[0] = 0u;
    nf->taken = 0u;
    if (bits >= NF_INF_BITS) {
        nf->done = 1;
        return false;
    }
    float delta = nf->delta;
    nf->total_arcs += nf->bucket_arcs;
    if (nf->total_arcs * 5 > nf->m * 3) {
        delta = 1e30f;   // fewer than 40 % of the arcs are left: one last bucket (the reference's order on the remainder)
    } else if (nf->bucket_arcs * 32 < nf->m) {
        delta *= nf->peaked ? 4.f : 2.f;
        if (!(delta < 1e30f)) delta = 1e30f;
    } else if (nf->bucket_arcs * 4 > nf->m) {
        nf->peaked = 1;
        delta = nf->delta0;
    }
    const float lo = __uint_as_float(bits);
    float hi = lo + delta;
    if (!(hi > lo)) hi = __uint_as_float(bits + 1u);   // width below one ulp of lo: the bucket is {lo}
    if (!(hi < 3.402823466e+38f)) hi = 3.402823466e+38f;   // never past "unreached" (FLT_MAX): the take pass must not pick those up
    nf->delta = delta;
    nf->lo = lo;
    nf->cutoff = hi;
    nf->bucket_arcs = 0;
    nf->buckets += 1;
    return true;
}

// Hook: what the surrounding level loop does around the passes (host loop: nothing; graph loop: level_loop.cu).
//   bool enabled() const          -- the pass has work in this launch
//   void trace(unsigned id) const -- B200_LOOP_TRACE timeline entry
//   void pre(NearFar *)           -- before the bucket is chosen (one thread)
//   void on_done(NearFar *)       -- no pending vertex is left
//   int *out() const              -- where the near frontier goes
//   void on_taken(NearFar *, uint32_t count)
template <class Hook>
__global__ void __launch_bounds__(NF_NT) sssp_pending_min_kernel(const float *__restrict__ dist, uint32_t n, NearFar *nf, Hook hook) {
    __shared__ unsigned s_min[NF_NT / 32];
    __shared__ bool s_last;
    if (!hook.enabled()) return;
    hook.trace(40);
    const float cut = nf->cutoff;
    unsigned best = NF_INF_BITS;
    const uint32_t tid = blockIdx.x * NF_NT + threadIdx.x, stride = gridDim.x * NF_NT;
    if ((reinterpret_cast<uintptr_t>(dist) & 15) == 0) {
        const float4 *d4 = reinterpret_cast<const float4 *>(dist);
        const uint32_t n4 = n >> 2;
        for (uint32_t i = tid; i < n4; i += stride) {
            const float4 d = __ldcg(d4 + i);
            if (d.x >= cut) best = min(best, __float_as_uint(d.x));
            if (d.y >= cut) best = min(best, __float_as_uint(d.y));
            if (d.z >= cut) best = min(best, __float_as_uint(d.z));
            if (d.w >= cut) best = min(best, __float_as_uint(d.w));
        }
        for (uint32_t i = (n4 << 2) + tid; i < n; i += stride) {
            const float d = __ldcg(dist + i);
            if (d >= cut) best = min(best, __float_as_uint(d));
        }
    } else {
        for (uint32_t i = tid; i < n; i += stride) {
            const float d = __ldcg(dist + i);
            if (d >= cut) best = min(best, __float_as_uint(d));
        }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < NF_NT / 32; ++w) best = min(best, s_min[w]);
        if (best < NF_INF_BITS) atomicMin(&nf->pending_min, best);
        __threadfence();
        s_last = atomicAdd(&nf->ticket[0], 1u) == gridDim.x - 1;
        if (s_last) {
            __threadfence();
            hook.pre(nf);
            if (!nf_open_bucket(nf)) hook.on_done(nf);
        }
    }
}

template <class Hook>
__global__ void __launch_bounds__(NF_NT) sssp_take_kernel(const float *__restrict__ dist, uint32_t n, NearFar *nf, Hook hook) {
    __shared__ uint32_t s_warp[NF_NT / 32];
    __shared__ uint32_t s_base;
    if (!hook.enabled() || nf->done) return;
    hook.trace(41);
    const float lo = nf->lo, hi = nf->cutoff;
    int *out = hook.out();
    const bool vec = (reinterpret_cast<uintptr_t>(dist) & 15) == 0;
    const uint32_t n4 = (n + 3) >> 2;   // groups of 4 consecutive vertices, one per thread per round
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t g0 = blockIdx.x * NF_NT; g0 < n4; g0 += gridDim.x * NF_NT) {   // (uniform per CTA: barriers inside)
        const uint32_t g = g0 + threadIdx.x, v0 = g << 2;
        float d[4] = {-1.f, -1.f, -1.f, -1.f};
        if (g < n4) {
            if (vec && v0 + 4 <= n) {
                const float4 q = __ldcg(reinterpret_cast<const float4 *>(dist) + g);
                d[0] = q.x, d[1] = q.y, d[2] = q.z, d[3] = q.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (v0 + k < n) d[k] = __ldcg(dist + v0 + k);
            }
        }
        uint32_t hit = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) hit |= (d[k] >= lo && d[k] < hi) ? 1u << k : 0u;
        const uint32_t cnt = __popc(hit);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0u;
#pragma unroll
            for (int w = 0; w < NF_NT / 32; ++w) {
                const uint32_t t = s_warp[w];
                s_warp[w] = tot;
                tot += t;
            }
            s_base = tot ? atomicAdd(&nf->taken, tot) : 0u;
        }
        __syncthreads();
        uint32_t pos = s_base + s_warp[warp] + incl - cnt;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (hit >> k & 1u) out[pos++] = (int)(v0 + k);
        __syncthreads();   // s_warp / s_base are rewritten by the next round
    }
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&nf->ticket[1], 1u) == gridDim.x - 1) {
            __threadfence();
            const uint32_t count = *reinterpret_cast<volatile unsigned *>(&nf->taken);
            nf->ticket[1] = 0u;
            hook.on_taken(nf, count);
        }
    }
}

// Host-driven loop (engine.cu): the host reads NearFar back after the two passes.
struct NearFarHostHook {
    int *out_list;
    long long bucket_arcs;   // counted by the host while the bucket was open
    __device__ __forceinline__ bool enabled() const { return true; }
    __device__ __forceinline__ void trace(unsigned) const {}
    __device__ __forceinline__ void pre(NearFar *nf) const { nf->bucket_arcs = bucket_arcs; }
    __device__ __forceinline__ void on_done(NearFar *) const {}
    __device__ __forceinline__ int *out() const { return out_list; }
    __device__ __forceinline__ void on_taken(NearFar *, uint32_t) const {}
};

// Mean of a strided sample of the weights (once per graph: the bucket width only steers the work, never the result).
static __global__ void __launch_bounds__(256) nf_weight_sample_kernel(const float *__restrict__ w, unsigned long long m, unsigned long long step,
                                                               unsigned int samples, double *sum) {
    __shared__ double s_sum[8];
    double acc = 0.0;
    for (unsigned int i = blockIdx.x * 256u + threadIdx.x; i < samples; i += gridDim.x * 256u) {
        const unsigned long long at = (unsigned long long)i * step;
        if (at < m) acc += (double)w[at];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) acc += s_sum[k];
        atomicAdd(sum, acc);
    }
}

}  // namespace b200
