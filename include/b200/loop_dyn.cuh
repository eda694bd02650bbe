// loop_dyn.cuh -- device-resident launch arguments of the graph-driven level loop.
//
// The reference's enactors branch on a count read back from the device after every operator
// (bfs_enactor.hxx:54-115: two blocking read-backs and ~10 launches per level).  Here the whole
// traversal can be ONE CUDA graph whose WHILE / IF conditional nodes are steered from the device
// (level_loop.cu), so the arguments that change from level to level -- which ping-pong buffer is
// the frontier, its length, the label to write, the look-back tag of the scan -- live in this
// struct in device memory instead of in kernel parameters.  Kernels take a nullable pointer to
// it; NULL means "use the by-value arguments" (the host-driven loop and the operator API).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace b200 {

struct LoopDyn {
    const int *in;       // current frontier list
    int *out;            // next frontier list
    uint32_t len;        // |in|
    uint32_t epoch;      // look-back tag of this level's scan (the bitmap -> list compaction uses epoch + 1)
    int next_label;      // level + 1
    uint32_t bsel;       // pull levels: which of the two frontier bitmaps is current
    uint32_t run;        // LOOP_RUN_* bits: which kernels of the loop body have work (the others return at once)
    // work-creating advance (NULL when the loop re-scans every level): this level's / the next level's quad scan and
    // row bounds, swapped by the decide step like in / out
    uint32_t *scanned_in;
    uint2 *rows_in;
    uint32_t *scanned_out;
    uint2 *rows_out;
    unsigned long long *trace;   // debug aid (B200_LOOP_TRACE=1): trace[0] = entries used, then (globaltimer ns << 8 | kernel id)
    uint32_t trace_cap;
};

// Kernel-entry timestamps of the graph-driven loop: with no host between the levels there is nothing to hang CUDA
// events on, so the kernels log their own start times (one thread, one atomic) when a trace buffer is attached.
__device__ __forceinline__ void loop_trace(const LoopDyn *dyn, unsigned id) {
#ifdef __CUDA_ARCH__
    if (dyn->trace && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned long long k = atomicAdd(dyn->trace, 1ull);
        if (k + 1 < dyn->trace_cap) dyn->trace[k + 1] = (t << 8) | id;
    }
#endif
}

// IF / SWITCH conditional nodes cost ~7 / ~4 us each on a B200 (profiles/microbench/graph_cond.cu) against
// ~1.1 us for a kernel that returns immediately, so the loop body is a flat kernel sequence and every
// kernel checks its bit.
enum : uint32_t { LOOP_RUN_PUSH = 1u, LOOP_RUN_PULL = 2u, LOOP_RUN_TO_PULL = 4u, LOOP_RUN_TO_PUSH = 8u,
                  LOOP_RUN_SCAN = 16u,     // the push level's frontier has no scan yet (first level, after a hand-over)
                  LOOP_RUN_SMALL = 32u,    // multi-GPU: the next push levels are small -- one persistent kernel runs them (p2p_bfs.cu)
                  LOOP_RUN_TAKE = 64u };   // near-far SSSP: the near frontier ran dry -- open the next bucket (near_far.cuh)

// Device-resident state of the near-far SSSP loop (b200_sssp_run).  The reference's enactor is
// frontier Bellman-Ford (sssp_enactor.hxx:40-72): every improved vertex is expanded in the next iteration, however
// far from final its distance is.  Here a vertex improved to a distance >= cutoff is NOT appended to the next
// frontier; it stays pending (its distance array entry is the only record) until the near frontier runs dry, then
// every vertex with a distance in the next bucket [lo, lo + delta) -- none of which can have been expanded yet,
// weights being non-negative -- becomes the near frontier.  Distances are the same fixed point either way.
struct NearFar {
    float cutoff;                // an improved vertex joins the next near frontier iff its new distance is below this
    float lo;                    // the open bucket is [lo, cutoff)
    float delta, delta0;         // bucket width now / at the start (widened once buckets get small, see near_far.cuh)
    unsigned int pending_min;    // bits of the smallest finite distance >= cutoff (distances are >= 0: bits order as floats)
    unsigned int taken;          // vertices the take pass moved into the near frontier
    unsigned int ticket[2];      // last-CTA-done tickets of the two passes
    long long bucket_arcs;       // arcs expanded since the bucket was opened
    long long total_arcs;        // ... in the buckets closed so far
    long long m;
    int buckets, done, peaked, pad;
};

__device__ __forceinline__ void near_far_reset(NearFar *nf, float delta0, long long m) {
    nf->cutoff = delta0;         // the first bucket is [0, delta0): the source is its only member
    nf->lo = 0.f;
    nf->delta = nf->delta0 = delta0;
    nf->pending_min = 0x7f7fffffu;   // FLT_MAX: nothing pending
    nf->taken = 0u;
    nf->ticket[0] = nf->ticket[1] = 0u;
    nf->bucket_arcs = 0;
    nf->total_arcs = 0;
    nf->m = m;
    nf->buckets = 1;
    nf->done = 0;
    nf->peaked = 0;
}

}  // namespace b200
