// p2p_bfs.cu -- multi-GPU BFS whose frontier exchange is FUSED INTO THE KERNELS over NVLink peer
// memory (no NCCL call on the data path).  New (the reference is single-GPU, README.md:4); the
// NCCL form of the same algorithm is multi_gpu.cu + mini_b200/dist.py and stays as the baseline.
//
// One process drives one GPU.  Every rank allocates one "symmetric heap" (same layout on every
// rank), exports it with cudaIpcGetMemHandle and maps the heaps of its peers, so a kernel can
// store straight into a peer's HBM through NVLink / NVSwitch:
//
//   push level : quad_advance_kernel<OUT_ROUTED> buckets accepted vertices by owner in its flush
//                and writes them DIRECTLY into the owner's inbox segment (box[p] is a peer
//                pointer) -- this is the alltoallv; a 1-warp kernel then publishes the per-peer
//                counts and runs a flag barrier across the GPUs; p2p_absorb_kernel applies
//                label-if-unvisited to what arrived (the uniquify filter of the exchange);
//   pull level : p2p_gather_or_kernel reads every peer's frontier-bitmap slice through NVLink into
//                the local full bitmap and folds it into `known` in the same pass -- this is the
//                allgather; the early-exit pull then runs over the local rows;
//   every level: p2p_stats_kernel writes this rank's {|F_next|, arcs, deg(F_next), sent} row into
//                every peer's heap, barriers, sums the rows (the allreduce) and leaves the result in
//                mapped pinned memory: ONE host synchronisation per level, zero collectives.
//
// Heap hazards are excluded by construction: slices are double-buffered (a level reads
// slice[sb] and writes slice[sb^1]); stats rows are double-buffered by barrier parity; inbox
// segments and counts are rewritten only after the level-closing barrier, which every rank
// reaches after its own absorb.  Barrier waits carry a wall-clock timeout (B200_ERR_TIMEOUT)
// so a lost peer cannot hang the GPU.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include "b200/operators.cuh"
#include "engine.cuh"
#include "graph_util.cuh"

using namespace b200;

namespace {

constexpr int P2P_MAX = MAX_DEST;
constexpr size_t P2P_CTRL_BYTES = 4096;
constexpr unsigned long long P2P_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
constexpr int P2P_TIMED_LEVELS = 64;

struct Ctrl {                                  // heap offset 0 on every rank
    unsigned long long flags[P2P_MAX];         // flags[p]: last barrier epoch rank p signalled to me
    unsigned long long counts[P2P_MAX];        // counts[p]: vertices rank p left in my inbox segment p
    unsigned long long stats[2][P2P_MAX][8];   // stats[parity][p]: rank p's row of the level summary
    unsigned long long hub_degree;             // level 0: degree of the source's row, told to every rank by its owner
};
static_assert(sizeof(Ctrl) <= P2P_CTRL_BYTES, "control block");

struct Peers {
    unsigned char *base[P2P_MAX];
};

enum { RES_SUM = 0, RES_OWN = 8, RES_TIMEOUT = 16, RES_WORDS = 32 };
enum { ROW_NEXT = 0, ROW_ARCS = 1, ROW_DEG = 2, ROW_SENT = 3, ROW_OVERFLOW = 4 };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spin loops poll the flag, not the clock: %globaltimer is read once every 256 polls (a read per poll made a flag
// barrier 7-9 us and a 148-CTA grid barrier 7 us in the traces of rounds 1 / 2).  Returns true once `limit_ns` have passed.
__device__ __forceinline__ bool spin_expired(uint32_t &polls, unsigned long long &t0, unsigned long long limit_ns) {
    if ((++polls & 255u) != 0u) return false;
    const unsigned long long now = globaltimer_ns();
    if (t0 == 0ull) {
        t0 = now;
        return false;
    }
    return now - t0 > limit_ns;
}

// One warp: lane p signals rank p (release: everything this rank wrote before -- in this kernel
// or in earlier kernels of the stream -- is visible to whoever acquires the flag) and waits for
// rank p's signal.  Returns false on every lane if any wait timed out.
__device__ __forceinline__ bool p2p_signal_wait(const Peers &peers, int me, int P, unsigned long long epoch) {
    const int p = (int)lane_id();
    bool ok = true;
    __threadfence_system();
    if (p < P && p != me) {
        st_release_sys(&reinterpret_cast<Ctrl *>(peers.base[p])->flags[me], epoch);
        const unsigned long long *mine = &reinterpret_cast<const Ctrl *>(peers.base[me])->flags[p];
        unsigned long long t0 = 0ull;
        uint32_t polls = 0u;
        while (ld_acquire_sys(mine) < epoch) {
            if (spin_expired(polls, t0, P2P_TIMEOUT_NS)) {
                ok = false;
                break;
            }
        }
    }
    ok = __all_sync(FULL_MASK, ok);
    __syncwarp();
    __threadfence_system();
    return ok;
}

__global__ void p2p_init_kernel(int32_t *labels, uint32_t *known, int32_t *frontier, int src, Partition part) {
    const uint32_t b = part.bit((uint32_t)src);
    known[b >> 5] |= 1u << (b & 31);              // every rank knows the source is visited
    if (part.owner((uint32_t)src) == part.me) {
        labels[part.row((uint32_t)src)] = 0;
        frontier[0] = src;
    }
}

// After a push advance: tell every peer how many vertices this rank left in its inbox, then barrier.
__global__ void __launch_bounds__(32) p2p_publish_counts_kernel(Peers peers, int me, int P, unsigned long long epoch,
                                                                const unsigned long long *box_counts,
                                                                unsigned long long *res) {
    const int p = (int)lane_id();
    if (p < P && p != me) st_relaxed_sys(&reinterpret_cast<Ctrl *>(peers.base[p])->counts[me], box_counts[p]);
    const bool ok = p2p_signal_wait(peers, me, P, epoch);
    if (p == 0 && !ok) res[RES_TIMEOUT] = 1ull;
}

// Level summary: this rank's row goes into every peer's heap, barrier, sum of the rows -> mapped
// pinned memory (what an allreduce + a D2H copy would have delivered).
__global__ void __launch_bounds__(32) p2p_stats_kernel(Peers peers, int me, int P, unsigned long long epoch, int parity,
                                                       const unsigned long long *counters,
                                                       const unsigned long long *box_counts, int push, int arcs_slot,
                                                       unsigned long long *res) {
    const int p = (int)lane_id();
    unsigned long long row[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) row[k] = 0;
    if (push) {
        row[ROW_NEXT] = box_counts[me];
        for (int q = 0; q < P; ++q)
            if (q != me) row[ROW_SENT] += box_counts[q];
    } else {
        row[ROW_NEXT] = counters[B200_CNT_OUT];
    }
    row[ROW_ARCS] = counters[arcs_slot];
    row[ROW_DEG] = counters[B200_CNT_AUX];
    row[ROW_OVERFLOW] = counters[B200_CNT_OVERFLOW];
    if (p < P) {
        unsigned long long *dst = reinterpret_cast<Ctrl *>(peers.base[p])->stats[parity][me];
#pragma unroll
        for (int k = 0; k < 8; ++k) st_relaxed_sys(dst + k, row[k]);
    }
    const bool ok = p2p_signal_wait(peers, me, P, epoch);
    if (p < 8) {
        const Ctrl *mine = reinterpret_cast<const Ctrl *>(peers.base[me]);
        unsigned long long s = 0;
        for (int q = 0; q < P; ++q) s += ld_relaxed_sys(&mine->stats[parity][q][p]);
        res[RES_SUM + p] = s;
        res[RES_OWN + p] = row[p];
    }
    if (p == 0 && !ok) res[RES_TIMEOUT] = 1ull;
    __threadfence_system();
}

// Receiver side of the push exchange over ALL inbox segments in one launch (counts live on the
// device): label-if-unvisited, survivors join the next frontier after the local discoveries.
__device__ __forceinline__ void p2p_absorb_body(const int *__restrict__ inbox, size_t seg_stride,
                                                const unsigned long long *counts, int me, int P,
                                                uint32_t *known, int *labels, int next_label, Partition part,
                                                int *next_frontier, unsigned long long capacity,
                                                unsigned long long *next_count, unsigned long long *counters,
                                                const uint32_t *__restrict__ offsets) {
    const unsigned lane = lane_id();
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long deg_sum = 0;
    for (int q = 0; q < P; ++q) {
        if (q == me) continue;
        const unsigned long long cnt = ld_relaxed_sys(counts + q);
        const int *seg = inbox + (size_t)q * seg_stride;
        for (unsigned long long base = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; base < cnt;
             base += nthreads) {
            const unsigned long long i = base + lane;
            bool fresh = false;
            int u = -1;
            if (i < cnt) {
                u = __ldcg(seg + i);
                const uint32_t b = part.bit((uint32_t)u);
                const uint32_t bit = 1u << (b & 31);
                fresh = !(known[b >> 5] & bit) && !(atomicOr(known + (b >> 5), bit) & bit);
                if (fresh) labels[part.row((uint32_t)u)] = next_label;
            }
            const unsigned mask = __ballot_sync(FULL_MASK, fresh);
            if (mask) {
                unsigned long long pos0 = 0;
                const unsigned leader = __ffs(mask) - 1;
                if (lane == leader) pos0 = atomicAdd(next_count, (unsigned long long)__popc(mask));
                pos0 = __shfl_sync(FULL_MASK, pos0, leader);
                if (fresh) {
                    const unsigned long long pos = pos0 + __popc(mask & lanemask_lt());
                    if (pos < capacity) next_frontier[pos] = u;
                    else counters[B200_CNT_OVERFLOW] = 1ull;
                    const uint32_t r = part.row((uint32_t)u);
                    deg_sum += __ldg(offsets + r + 1) - __ldg(offsets + r);
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) deg_sum += __shfl_xor_sync(FULL_MASK, deg_sum, d);
    if (lane == 0 && deg_sum) atomicAdd(&counters[B200_CNT_AUX], deg_sum);
}
__global__ void __launch_bounds__(256) p2p_absorb_kernel(const int *__restrict__ inbox, size_t seg_stride,
                                                         const unsigned long long *counts, int me, int P,
                                                         uint32_t *known, int *labels, int next_label, Partition part,
                                                         int *next_frontier, unsigned long long capacity,
                                                         unsigned long long *next_count, unsigned long long *counters,
                                                         const uint32_t *__restrict__ offsets) {
    p2p_absorb_body(inbox, seg_stride, counts, me, P, known, labels, next_label, part, next_frontier, capacity, next_count,
                    counters, offsets);
}

// next frontier list (length on the device) -> this rank's bitmap slice (pre-cleared)
__global__ void p2p_list_to_slice_kernel(const int *__restrict__ list, const unsigned long long *len_ptr, uint32_t *slice,
                                         Partition part) {
    const unsigned long long len = *len_ptr;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const uint32_t r = part.row((uint32_t)list[i]);
        atomicOr(slice + (r >> 5), 1u << (r & 31));
    }
}

// The allgather of a pull level, fused with the fold into `known`: read every rank's frontier
// slice through NVLink (own slice locally), write the full rank-major bitmap, OR into known.
__global__ void __launch_bounds__(256) p2p_gather_or_kernel(Peers peers, size_t slice_off, uint32_t quads_per_slice, int P,
                                                            uint4 *__restrict__ full, uint4 *__restrict__ known) {
    const size_t total = (size_t)quads_per_slice * (size_t)P;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t p = (uint32_t)(i / quads_per_slice);
        const uint32_t j = (uint32_t)(i - (size_t)p * quads_per_slice);
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(peers.base[p] + slice_off) + j);
        full[i] = v;
        if (v.x | v.y | v.z | v.w) {
            uint4 k = known[i];
            k.x |= v.x; k.y |= v.y; k.z |= v.z; k.w |= v.w;
            known[i] = k;
        }
    }
}

struct PartItem {   // local row r of this rank -> its global vertex id
    Partition part;
    __device__ __forceinline__ int operator()(uint32_t idx) const { return (int)part.global_id(part.me, idx); }
};


// ============================================================================================
// Graph-driven level loop (see level_loop.cu for the single-GPU form and the measurements behind
// it): the whole multi-GPU traversal is ONE CUDA graph per rank -- prologue, then a WHILE
// conditional node whose flat body holds every kernel of a level; each kernel returns at once
// unless its LOOP_RUN_* bit is set.  All ranks compute the same global level summary in
// p2p_stats_decide_kernel (the in-heap allreduce) and apply the same push / pull / stop rules to
// it on the device, so no rank returns to its host between levels: zero host synchronisations and
// zero collectives per level.  Barrier epochs, the stats parity and the look-back tags live in
// P2PLoopState instead of kernel parameters.
// ============================================================================================
struct P2PLoopParams {      // mapped pinned, read by the init kernel
    int32_t src, mode;
    float alpha, beta;
    long long m_global;
    unsigned long long bar_epoch0;
    uint32_t lb_epoch0, stats_seq0;
    unsigned long long *trace;
    uint32_t trace_cap, small_on;
    long long small_arcs, small_verts;   // a push level runs in the persistent small-level kernel while its frontier is this small
    long long hub_min;                   // level 0: a source row of at least this many arcs goes through the bitmap exchange
};
struct P2PLevelRec {
    int32_t direction, exchange;   // exchange: 0 = vertex ids through the inboxes, 1 = bitmap slices
    long long frontier_len, arcs, discovered, sent;
};
struct P2PLoopResult {      // mapped pinned, written by the decide step
    int32_t status, num_levels;
    long long reached, total_arcs, launches;
    unsigned long long bar_epoch;
    uint32_t stats_seq, pad;
    P2PLevelRec level[B200_MAX_LEVELS];
};
struct P2PLoopState {       // device
    LoopDyn dyn;            // in / out / len = this rank's frontier list; bsel = which slice holds the current frontier
    unsigned long long bar_epoch;
    uint32_t stats_seq, timeout;
    int32_t level, pull, mode, kernels_per_level;
    float alpha, beta;
    long long n, m_unexplored, flen, reached, total_arcs, launches;
    long long small_arcs, small_verts, hub_min;
    uint32_t small_on;
    int32_t src;
};

// ---- persistent small-level kernel: shared state -------------------------------------------
constexpr int SMALL_NT = 512;
constexpr int SMALL_SLOTS = 4;
constexpr uint32_t SMALL_FLUSH = 256; // ids per flush of a staging row: ONE slot reservation (a same-address atomic) per 256 winners
constexpr uint32_t SMALL_STAGE = SMALL_FLUSH + 32;   // slots per (warp, destination) row: < SMALL_FLUSH left after a flush + <= 32 new
constexpr size_t SMALL_SMEM = sizeof(int) * (SMALL_NT / 32) * P2P_MAX * SMALL_STAGE;   // 147 KB of dynamic shared memory
constexpr uint32_t SMALL_ROW = 512;   // rows up to this many arcs: one warp; longer rows: pieces of SMALL_ROW arcs over the whole grid
constexpr unsigned long long SMALL_GRID_TIMEOUT_NS = 5ull * 1000ull * 1000ull * 1000ull;
struct SmallCounters {      // one slot per level (level & 3): nobody has to wait for a reset
    unsigned long long next_cnt, arcs, deg, big_cnt, overflow, pad;
    unsigned long long send_cnt[P2P_MAX];
};
struct PullCounters {       // the pull-levels kernel's per-level counters (indexed by B200_CNT_*), one slot per level & 3
    unsigned long long c[B200_NUM_COUNTERS];
};
struct SmallShared {        // device memory
    SmallCounters slot[SMALL_SLOTS];
    unsigned int bar_count, bar_gen;   // grid barrier of the small-level / pull-levels kernels
    unsigned int abort;
    unsigned int absorb_tile;          // dynamic tile claims of p2p_absorb_bits_kernel (re-armed by the decide kernel that follows it)
};

struct FrontierQuadsDynPart {   // FrontierQuads over the device-selected list, rows of a 1D partition
    const LoopDyn *dyn;
    const uint32_t *offsets;
    uint2 *rows;
    uint32_t row_shift;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const int v = dyn->in[i];
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

__device__ __forceinline__ void small_zero_slot(SmallCounters *c) {
    c->next_cnt = 0ull; c->arcs = 0ull; c->deg = 0ull; c->big_cnt = 0ull; c->overflow = 0ull;
    for (int i = 0; i < P2P_MAX; ++i) c->send_cnt[i] = 0ull;
}

__global__ void p2p_loop_init_kernel(const P2PLoopParams *p, P2PLoopState *s, int32_t *labels, uint32_t *known, uint32_t *done,
                                     int32_t *f0, int32_t *f1, long long n, Partition part, unsigned long long *counters,
                                     unsigned int *tile_counters, SmallShared *sh, PullCounters *pull_slots,
                                     int kernels_per_level) {
    const int src = p->src;
    const uint32_t b = part.bit((uint32_t)src);
    known[b >> 5] |= 1u << (b & 31);
    const bool mine = part.owner((uint32_t)src) == part.me;
    if (mine) {
        const uint32_t r = part.row((uint32_t)src);
        labels[r] = 0;
        done[r >> 5] |= 1u << (r & 31);
        f0[0] = src;
    }
    s->dyn.in = f0;
    s->dyn.out = f1;
    s->dyn.len = mine ? 1u : 0u;
    s->dyn.epoch = p->lb_epoch0 & 0x3FFFFFFFu;
    s->dyn.next_label = 1;
    s->dyn.bsel = 0u;
    s->dyn.run = p->small_on ? LOOP_RUN_SMALL : LOOP_RUN_PUSH;   // level 0 is one vertex
    s->dyn.scanned_in = nullptr;   // big push levels scan their frontier (it comes out of a bitmap)
    s->dyn.rows_in = nullptr;
    s->dyn.scanned_out = nullptr;
    s->dyn.rows_out = nullptr;
    s->dyn.trace = p->trace;
    s->dyn.trace_cap = p->trace_cap;
    if (p->trace) p->trace[0] = 0ull;
    s->bar_epoch = p->bar_epoch0;
    s->stats_seq = p->stats_seq0;
    s->timeout = 0u;
    s->level = 0;
    s->pull = 0;
    s->mode = p->mode;
    s->kernels_per_level = kernels_per_level;
    s->alpha = p->alpha;
    s->beta = p->beta;
    s->n = n;
    s->m_unexplored = p->m_global;
    s->flen = 1;
    s->reached = 1;
    s->total_arcs = 0;
    s->launches = 1;
    s->small_arcs = p->small_arcs;
    s->small_verts = p->small_verts;
    s->small_on = p->small_on;
    s->hub_min = p->hub_min;
    s->src = p->src;
    for (int i = 0; i < B200_NUM_COUNTERS; ++i) counters[i] = 0ull;
    tile_counters[0] = 0u;
    tile_counters[1] = 0u;
    for (int i = 0; i < SMALL_SLOTS; ++i) {
        small_zero_slot(&sh->slot[i]);
        for (int k = 0; k < B200_NUM_COUNTERS; ++k) pull_slots[i].c[k] = 0ull;
    }
    sh->bar_count = 0u;
    sh->abort = 0u;
    sh->absorb_tile = 0u;
}

// ---------------------------------------------------------------------------------------------
// The small levels of a traversal (level 0: one vertex; the tail after the pull phase; every level of a
// small graph) cost a CUDA-graph iteration of ~10 kernel nodes and two flag barriers each in the flat body
// -- 58 us per level at 8 GPUs for levels that move a few thousand vertices (round-1 trace).  This kernel
// runs CONSECUTIVE small push levels inside one launch: per level two grid barriers and two cross-GPU flag
// barriers, no kernel boundary.  One CTA per SM, all resident (the host sizes the grid from the occupancy).
//   phase A  push over the local frontier: rows <= SMALL_ROW arcs one warp each, longer rows (the hub of
//            level 0: 1.7 M arcs at scale 26) cut into SMALL_ROW-arc pieces dealt over every warp of the grid;
//            a won vertex is labelled in place (owned) or stored into its owner's inbox over NVLink;
//   barrier  grid, then flags to / from every peer (the counts travel with the flag);
//   phase B  absorb the inbox segments: label-if-unvisited, append to the next frontier;
//   barrier  grid, stats row to every peer, flags; every CTA sums the rows and takes the same decision.
// The level state lives in registers (identically in every CTA); CTA 0 writes it back on exit for the
// kernels that follow in the same graph iteration (pull or big push), so they run without another decide.
// Every wait has a wall-clock timeout that raises the abort flag: a lost peer or a CTA that never became
// resident ends the traversal with B200_ERR_TIMEOUT instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------
struct SmallArgs {
    Peers peers;
    int me, P;
    Partition part;
    const uint32_t *offsets;      // local CSR (global column ids)
    const int *indices;
    uint32_t *known, *done;
    int *labels;
    P2PLoopState *s;
    SmallShared *sh;
    uint32_t *big;                // frontier slots of the long rows of a level
    size_t off_inbox, off_slice0, off_slice1;
    uint32_t wl;                  // words per slice
    const uint32_t *iso;          // no-in-arc bitmap of the local rows (nullable: derived from pull_offsets)
    const uint32_t *pull_offsets;
    P2PLoopResult *res;
    cudaGraphConditionalHandle h_while;
    PullCounters *pull_slots;     // re-armed on exit (the pull-levels kernel may run next; it re-arms this kernel's)
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Whole grid.  s_gen: this CTA's copy of the barrier generation (shared memory).  Returns false once aborted.
// sys = true: what this CTA stored into PEER memory is ordered before whatever a thread signals after the barrier --
// one system-scope fence by thread 0 after the CTA barrier (cumulative over the CTA's stores), not one per thread
// (75 K MEMBAR.SYS per level made a small level 15 us longer in the 2-GPU trace).
__device__ __forceinline__ bool small_grid_barrier(SmallShared *sh, unsigned *s_gen, bool sys = false) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned g = *s_gen;
        if (sys) __threadfence_system();
        else __threadfence();
        if (atomicAdd(&sh->bar_count, 1u) == gridDim.x - 1u) {
            sh->bar_count = 0u;
            __threadfence();
            atomicAdd(&sh->bar_gen, 1u);
        } else {
            unsigned long long t0 = 0ull;
            uint32_t polls = 0u;
            while (ld_acquire_gpu_u32(&sh->bar_gen) == g) {
                if (spin_expired(polls, t0, SMALL_GRID_TIMEOUT_NS)) {
                    atomicExch(&sh->abort, 1u);
                    break;
                }
                if ((polls & 63u) == 0u && ld_relaxed_gpu_u32(&sh->abort)) break;
            }
        }
        *s_gen = g + 1u;
    }
    __syncthreads();
    return ld_relaxed_gpu_u32(&sh->abort) == 0u;
}

// Every CTA waits for the flags of all peers (they live in this rank's own heap: polling is local).
__device__ __forceinline__ bool small_wait_peers(const SmallArgs &a, unsigned long long epoch) {
    if (threadIdx.x < (unsigned)a.P && (int)threadIdx.x != a.me) {
        const unsigned long long *mine = &reinterpret_cast<const Ctrl *>(a.peers.base[a.me])->flags[threadIdx.x];
        unsigned long long t0 = 0ull;
        uint32_t polls = 0u;
        while (ld_acquire_sys(mine) < epoch) {
            if (spin_expired(polls, t0, P2P_TIMEOUT_NS)) {
                atomicExch(&a.sh->abort, 1u);
                break;
            }
            if ((polls & 63u) == 0u && ld_relaxed_gpu_u32(&a.sh->abort)) break;
        }
    }
    __syncthreads();
    return ld_relaxed_gpu_u32(&a.sh->abort) == 0u;
}

__global__ void __launch_bounds__(SMALL_NT, 1) p2p_small_levels_kernel(SmallArgs a) {
    constexpr int NW = SMALL_NT / 32, U = 8;
    extern __shared__ int s_stage_dyn[];     // [NW][P2P_MAX][SMALL_STAGE]
    __shared__ uint32_t s_fill[NW][P2P_MAX];
    __shared__ unsigned s_gen;
    __shared__ unsigned long long s_sum[8];
    P2PLoopState *s = a.s;
    loop_trace(&s->dyn, 20);
    if (!(s->dyn.run & LOOP_RUN_SMALL)) return;
    SmallShared *sh = a.sh;
    if (ld_relaxed_gpu_u32(&sh->abort)) return;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5, lt_mask = lanemask_lt();
    const uint32_t gwarp = blockIdx.x * NW + warp, total_warps = gridDim.x * NW;
    const int me = a.me, P = a.P;
    const Partition part = a.part;
    Ctrl *my_ctrl = reinterpret_cast<Ctrl *>(a.peers.base[me]);
    const int *my_inbox = reinterpret_cast<const int *>(a.peers.base[me] + a.off_inbox);
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    if (threadIdx.x == 0) s_gen = ld_relaxed_gpu_u32(&sh->bar_gen);
    int *wbuf = s_stage_dyn + (size_t)warp * P2P_MAX * SMALL_STAGE;
    uint32_t *wfill = s_fill[warp];
    if (lane < P2P_MAX) wfill[lane] = 0u;

    // the level state, identically in every CTA
    int level = s->level;
    const int mode = s->mode;
    const float alpha = s->alpha;
    long long m_unexplored = s->m_unexplored, flen = s->flen, reached = s->reached, total_arcs = s->total_arcs;
    const long long small_arcs = s->small_arcs;
    unsigned long long bar_epoch = s->bar_epoch;
    uint32_t stats_seq = s->stats_seq, bsel = s->dyn.bsel;
    const int *in = s->dyn.in;
    int *out = s->dyn.out;
    uint32_t len = s->dyn.len;
    long long launches = s->launches;
    int status = B200_OK;
    uint32_t next_run = 0u;
    bool finished = false;
    __syncthreads();

    // ---- level 0 from a hub: one row (1.7 M arcs at scale 26) expanded by this kernel's id-by-id path cost 55 us on the
    // owner at 8 GPUs plus 21 us of absorbing and 20 us of building the first pull level's bitmap slice (dealing the row
    // out to the peers first did not help: 32 us of copying, and the expansion stayed latency-bound).  Such a level is a
    // BIG level in everything but its vertex count: the owner tells every rank the degree of the source (one flag
    // barrier), and from `hub_min` arcs on the level goes through the bitmap exchange -- the quad advance spreads the one
    // row over the whole GPU and only claims bits, and every rank picks its slice out of the owner's `known`.
    bool hub_level0 = false;
    if (level == 0 && P > 1) {
        const uint32_t src = (uint32_t)s->src;
        const int owner0 = (int)part.owner(src);
        unsigned long long hub_deg = 0ull;
        if (owner0 == me) {
            const uint32_t r = part.row(src);
            hub_deg = __ldg(a.offsets + r + 1) - __ldg(a.offsets + r);
        }
        ++bar_epoch;
        if (blockIdx.x == 0 && threadIdx.x < (unsigned)P && (int)threadIdx.x != me) {
            Ctrl *pc = reinterpret_cast<Ctrl *>(a.peers.base[threadIdx.x]);
            if (owner0 == me) st_relaxed_sys(&pc->hub_degree, hub_deg);
            st_release_sys(&pc->flags[me], bar_epoch);
        }
        const bool ok0 = small_wait_peers(a, bar_epoch);
        if (owner0 != me) hub_deg = ok0 ? ld_relaxed_sys(&my_ctrl->hub_degree) : 0ull;
        hub_level0 = ok0 && (long long)hub_deg >= s->hub_min;
    }

    for (;;) {
        if (hub_level0) {
            next_run = LOOP_RUN_PUSH;    // level 0 through scan + claim-only advance + bitmap absorb, in this graph iteration
            break;
        }
        SmallCounters *c = &sh->slot[level & (SMALL_SLOTS - 1)];
        if (lead) small_zero_slot(&sh->slot[(level + 1) & (SMALL_SLOTS - 1)]);   // last used three levels ago
        const int next_label = level + 1;
        unsigned long long arc_cnt = 0, deg_sum = 0;

        // this warp's staging rows (see expand) and their flush: `count` ids of destination p leave for its box
        auto flush_row = [&](int p, uint32_t count) {
            unsigned long long base = 0ull;
            if (lane == 0) base = atomicAdd(p == me ? &c->next_cnt : &c->send_cnt[p], (unsigned long long)count);
            base = __shfl_sync(FULL_MASK, base, 0);
            int *box = p == me ? out : reinterpret_cast<int *>(a.peers.base[p] + a.off_inbox) + (size_t)me * part.n_local;
            for (uint32_t k = lane; k < count; k += 32u) {
                if (base + k < part.n_local) box[base + k] = wbuf[p * SMALL_STAGE + k];
                else c->overflow = 1ull;
            }
        };
        // the per-arc step of a warp over arcs [b, e) of one row piece, 32 * U arcs at a time: all index loads, then all
        // probes, then all claims are in flight together; the winners are counted per destination over the whole tile and
        // every destination's slots are reserved with ONE atomic, the P atomics issued side by side by lanes 0 .. P-1
        auto expand = [&](uint32_t b, uint32_t e) {
            for (uint32_t e0 = b; e0 < e; e0 += 32u * U) {
                int d[U], owner[U];
                uint32_t kb[U], w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t ee = e0 + 32u * u + lane;
                    d[u] = ee < e ? __ldg(a.indices + ee) : -1;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    kb[u] = 0u;
                    w[u] = 0xffffffffu;
                    if (d[u] >= 0) {
                        kb[u] = part.bit((uint32_t)d[u]);
                        w[u] = a.known[kb[u] >> 5];
                    }
                }
                uint32_t old[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    old[u] = 0xffffffffu;
                    if (!((w[u] >> (kb[u] & 31)) & 1u)) old[u] = atomicOr(a.known + (kb[u] >> 5), 1u << (kb[u] & 31));
                }
                bool any = false;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool won = !((old[u] >> (kb[u] & 31)) & 1u);
                    owner[u] = won ? (int)part.owner((uint32_t)d[u]) : -1;
                    any |= won;
                    if (won && owner[u] == me) {
                        const uint32_t r = part.row((uint32_t)d[u]);
                        atomicOr(a.done + (r >> 5), 1u << (r & 31));
                        a.labels[r] = next_label;
                        deg_sum += __ldg(a.offsets + r + 1) - __ldg(a.offsets + r);
                    }
                }
                if (!__any_sync(FULL_MASK, any)) continue;
                // winners go to their owner's box through this warp's staging rows (one row per destination in shared
                // memory): a row is flushed SMALL_FLUSH ids at a time -- ONE slot reservation and full 128-byte stores, local
                // or over NVLink.  (A reservation per 32 winners was 20 K atomics on one counter for the 0.64 M discoveries of a
                // scale-25 hub row: the warps stood in line for ~27 of the level's 51 us.)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (!__any_sync(FULL_MASK, owner[u] >= 0)) continue;
                    for (int p = 0; p < P; ++p) {
                        const unsigned mask = __ballot_sync(FULL_MASK, owner[u] == p);
                        if (!mask) continue;
                        uint32_t f = wfill[p];
                        if (owner[u] == p) wbuf[p * SMALL_STAGE + f + __popc(mask & lt_mask)] = d[u];
                        f += __popc(mask);
                        __syncwarp();
                        if (f >= SMALL_FLUSH) {
                            flush_row(p, SMALL_FLUSH);
                            const int x = lane < f - SMALL_FLUSH ? wbuf[p * SMALL_STAGE + SMALL_FLUSH + lane] : 0;   // (< 32 are left)
                            __syncwarp();
                            if (lane < f - SMALL_FLUSH) wbuf[p * SMALL_STAGE + lane] = x;
                            f -= SMALL_FLUSH;
                        }
                        __syncwarp();
                        if (lane == 0) wfill[p] = f;
                        __syncwarp();
                    }
                }
            }
            arc_cnt += e - b;   // (counted once per piece: same value on every lane)
        };

        // ---- phase A, pass 1: short rows, one warp each; long rows are queued
        loop_trace(&s->dyn, 21);
        for (uint32_t i = gwarp; i < len; i += total_warps) {
            // (the frontier lists and the long-row queue are rewritten level after level inside this one launch: L1
            // may hold a line of an earlier level, so they are read through L2)
            const uint32_t r = (uint32_t)__ldcg(in + i) >> part.log_p;
            const uint32_t b = __ldg(a.offsets + r), e = __ldg(a.offsets + r + 1);
            if (e - b > SMALL_ROW) {
                if (lane == 0) a.big[atomicAdd(&c->big_cnt, 1ull)] = r;
                continue;
            }
            if (e > b) expand(b, e);
        }
        for (int p = 0; p < P; ++p) {                // what is left in this warp's staging rows
            const uint32_t f = wfill[p];
            __syncwarp();
            if (f) flush_row(p, f);
            if (lane == 0) wfill[p] = 0u;
        }
        // (system scope: when no long row was queued this is already the "sends are out" barrier)
        loop_trace(&s->dyn, 26);
        if (!small_grid_barrier(sh, &s_gen, true)) break;
        loop_trace(&s->dyn, 27);
        // ---- pass 2: the pieces of the long rows, dealt round-robin over the warps of the grid
        const uint32_t nbig = (uint32_t)ld_volatile_u64(&c->big_cnt);
        if (nbig) {
            uint32_t skew = 0;
            for (uint32_t k = 0; k < nbig; ++k) {
                const uint32_t r = __ldcg(a.big + k);
                const uint32_t b = __ldg(a.offsets + r), e = __ldg(a.offsets + r + 1);
                const uint32_t pieces = (e - b + SMALL_ROW - 1) / SMALL_ROW;
                for (uint32_t p = (gwarp + total_warps - skew) % total_warps; p < pieces; p += total_warps) {
                    const uint32_t pb = b + p * SMALL_ROW;
                    expand(pb, e - pb > SMALL_ROW ? pb + SMALL_ROW : e);
                }
                skew = (skew + pieces) % total_warps;
            }
        }
        if (nbig) {
            for (int p = 0; p < P; ++p) {            // what is left in this warp's staging rows
                const uint32_t f = wfill[p];
                __syncwarp();
                if (f) flush_row(p, f);
                if (lane == 0) wfill[p] = 0u;
            }
            if (!small_grid_barrier(sh, &s_gen, true)) break;   // (system scope: the stores into the peers' inboxes, before the flag)
        }
        // ---- everybody's sends are out: counts + flag to every peer, then wait for theirs
        ++bar_epoch;
        if (blockIdx.x == 0 && threadIdx.x < (unsigned)P && (int)threadIdx.x != me) {
            Ctrl *pc = reinterpret_cast<Ctrl *>(a.peers.base[threadIdx.x]);
            st_relaxed_sys(&pc->counts[me], ld_volatile_u64(&c->send_cnt[threadIdx.x]));
            st_release_sys(&pc->flags[me], bar_epoch);
        }
        loop_trace(&s->dyn, 22);
        if (!small_wait_peers(a, bar_epoch)) break;
        loop_trace(&s->dyn, 23);
        // ---- phase B: absorb what the peers left in the inbox segments.  The P - 1 segments are walked as ONE list (a
        // loop over the peers ran their dependent load -> probe -> claim chains one after the other: 14 us at 8 GPUs for a
        // level that received 25 ids)
        {
            unsigned long long pre[P2P_MAX + 1];
            pre[0] = 0ull;
#pragma unroll
            for (int q = 0; q < P2P_MAX; ++q)
                pre[q + 1] = pre[q] + ((q < P && q != me) ? ld_relaxed_sys(&my_ctrl->counts[q]) : 0ull);
            const unsigned long long cnt = pre[P2P_MAX];
            auto entry = [&](unsigned long long i) -> int {   // the i-th id of the concatenated segments
                int q = 0;
#pragma unroll
                for (int k = 1; k < P2P_MAX; ++k) q += i >= pre[k];
                return __ldcg(my_inbox + (size_t)q * part.n_local + (i - pre[q]));
            };
            // a warp takes 32 * G entries at a time and reserves their frontier slots with ONE atomic (a same-address
            // atomic per 32 entries was what the absorb of level 0 -- 0.5 M ids -- spent its 30 us on)
            constexpr int G = 8;
            for (unsigned long long base = (unsigned long long)gwarp * (32ull * G); base < cnt; base += (unsigned long long)total_warps * (32ull * G)) {
                int u[G];
                unsigned mask[G];
                uint32_t tot = 0u, dw[G], old[G];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const unsigned long long i = base + 32ull * j + lane;
                    u[j] = i < cnt ? entry(i) : -1;
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {          // all probes of `done` in flight
                    dw[j] = 0xffffffffu;
                    if (u[j] >= 0) dw[j] = a.done[part.row((uint32_t)u[j]) >> 5];
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {          // then all claims
                    const uint32_t r = part.row((uint32_t)(u[j] >= 0 ? u[j] : 0)), bit = 1u << (r & 31);
                    old[j] = 0xffffffffu;
                    if (u[j] >= 0 && !(dw[j] & bit)) old[j] = atomicOr(a.done + (r >> 5), bit);
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const uint32_t r = part.row((uint32_t)(u[j] >= 0 ? u[j] : 0)), bit = 1u << (r & 31);
                    const bool fresh = !(old[j] & bit);
                    if (fresh) {
                        a.labels[r] = next_label;
                        atomicOr(a.known + (size_t)me * a.wl + (r >> 5), bit);
                        deg_sum += __ldg(a.offsets + r + 1) - __ldg(a.offsets + r);
                    } else {
                        u[j] = -1;
                    }
                    mask[j] = __ballot_sync(FULL_MASK, fresh);
                    tot += __popc(mask[j]);
                }
                if (tot == 0u) continue;
                unsigned long long pos0 = 0;
                if (lane == 0) pos0 = atomicAdd(&c->next_cnt, (unsigned long long)tot);
                pos0 = __shfl_sync(FULL_MASK, pos0, 0);
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if (u[j] >= 0) {
                        const unsigned long long pos = pos0 + __popc(mask[j] & lt_mask);
                        if (pos < part.n_local) out[pos] = u[j];
                        else c->overflow = 1ull;
                    }
                    pos0 += __popc(mask[j]);
                }
            }
        }
        // level totals of this CTA -> the slot
        loop_trace(&s->dyn, 28);
        if (threadIdx.x < 2) s_sum[threadIdx.x] = 0ull;
        __syncthreads();
#pragma unroll
        for (int d2 = 16; d2 > 0; d2 >>= 1) deg_sum += __shfl_xor_sync(FULL_MASK, deg_sum, d2);
        if (lane == 0) {
            if (arc_cnt) atomicAdd(&s_sum[0], arc_cnt);
            if (deg_sum) atomicAdd(&s_sum[1], deg_sum);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_sum[0]) atomicAdd(&c->arcs, s_sum[0]);
            if (s_sum[1]) atomicAdd(&c->deg, s_sum[1]);
        }
        if (!small_grid_barrier(sh, &s_gen)) break;
        loop_trace(&s->dyn, 29);
        // ---- level summary: this rank's row into every heap, flags, sum of the rows
        const unsigned long long next_local = ld_volatile_u64(&c->next_cnt);
        const int parity = (int)(stats_seq & 1u);
        ++bar_epoch;
        if (blockIdx.x == 0 && threadIdx.x < (unsigned)P) {
            unsigned long long sent = 0;
            for (int q = 0; q < P; ++q)
                if (q != me) sent += ld_volatile_u64(&c->send_cnt[q]);
            unsigned long long *dst = reinterpret_cast<Ctrl *>(a.peers.base[threadIdx.x])->stats[parity][me];
            st_relaxed_sys(dst + ROW_NEXT, next_local);
            st_relaxed_sys(dst + ROW_ARCS, ld_volatile_u64(&c->arcs));
            st_relaxed_sys(dst + ROW_DEG, ld_volatile_u64(&c->deg));
            st_relaxed_sys(dst + ROW_SENT, sent);
            st_relaxed_sys(dst + ROW_OVERFLOW, ld_volatile_u64(&c->overflow));
            if ((int)threadIdx.x != me) st_release_sys(&reinterpret_cast<Ctrl *>(a.peers.base[threadIdx.x])->flags[me], bar_epoch);
        }
        __syncthreads();                 // (CTA 0: its own row is written before its threads read it back)
        loop_trace(&s->dyn, 24);
        if (!small_wait_peers(a, bar_epoch)) break;
        loop_trace(&s->dyn, 25);
        if (threadIdx.x < 8) {
            // own row from the local counters (final since the grid barrier; CTA 0's copy in the heap may still be in
            // flight for the other CTAs), the peers' rows from the heap
            unsigned long long t = 0;
            if (threadIdx.x == ROW_NEXT) t = next_local;
            else if (threadIdx.x == ROW_ARCS) t = ld_volatile_u64(&c->arcs);
            else if (threadIdx.x == ROW_DEG) t = ld_volatile_u64(&c->deg);
            else if (threadIdx.x == ROW_OVERFLOW) t = ld_volatile_u64(&c->overflow);
            else if (threadIdx.x == ROW_SENT)
                for (int q = 0; q < P; ++q)
                    if (q != me) t += ld_volatile_u64(&c->send_cnt[q]);
            for (int q = 0; q < P; ++q)
                if (q != me) t += ld_relaxed_sys(&my_ctrl->stats[parity][q][threadIdx.x]);
            s_sum[threadIdx.x] = t;
        }
        __syncthreads();
        const long long found = (long long)s_sum[ROW_NEXT], arcs = (long long)s_sum[ROW_ARCS];
        const long long next_deg = (long long)s_sum[ROW_DEG], sent_all = (long long)s_sum[ROW_SENT];
        const bool overflow = s_sum[ROW_OVERFLOW] != 0ull;
        __syncthreads();                 // s_sum is reused by the next level
        // ---- the decision of p2p_stats_decide_kernel, in every CTA alike
        if (lead && level < B200_MAX_LEVELS) {
            P2PLevelRec *r = &a.res->level[level];
            r->direction = 0;
            r->exchange = 0;
            r->frontier_len = flen;
            r->arcs = arcs;
            r->discovered = found;
            r->sent = sent_all;
        }
        loop_trace(&s->dyn, 64u + ((unsigned)level & 63u));   // level `level` closed
        total_arcs += arcs;
        launches += 1;
        ++level;
        ++stats_seq;
        if (overflow) {
            status = B200_ERR_OVERFLOW;
            finished = true;
            break;
        }
        if (found == 0) {
            finished = true;
            break;
        }
        reached += found;
        m_unexplored -= arcs;
        {
            const int *t = in;
            in = out;
            out = const_cast<int *>(t);
        }
        len = (uint32_t)next_local;
        const bool to_pull = mode == B200_BFS_BEAMER && (double)next_deg > (double)m_unexplored / alpha && found > flen;
        flen = found;
        if (to_pull) {
            // the new frontier as this rank's bitmap slice (the first pull level gathers the slices), and the rows
            // without in-arcs count as done from here on (engine.cuh): pull levels skip them
            // (The slice is NOT cleared first: the prologue zeroes both slices, and whatever a big push level or an earlier
            // pull phase of THIS traversal left there are vertices of earlier levels, all of whose neighbours are visited --
            // as frontier bits they cannot label anybody.)
            // (The "no in-arc" preset of `done` is applied by the pull-levels kernel that follows: LOOP_RUN_TO_PULL.)
            uint32_t *slice_w = reinterpret_cast<uint32_t *>(a.peers.base[me] + (bsel ? a.off_slice0 : a.off_slice1));
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
                const uint32_t r = part.row((uint32_t)__ldcg(in + i));
                atomicOr(slice_w + (r >> 5), 1u << (r & 31));
            }
            if (!small_grid_barrier(sh, &s_gen, true)) break;
            ++bar_epoch;
            if (blockIdx.x == 0 && threadIdx.x < (unsigned)P && (int)threadIdx.x != me)
                st_release_sys(&reinterpret_cast<Ctrl *>(a.peers.base[threadIdx.x])->flags[me], bar_epoch);
            if (!small_wait_peers(a, bar_epoch)) break;
            bsel ^= 1u;
            next_run = LOOP_RUN_PULL | LOOP_RUN_TO_PULL;
            break;
        }
        if (next_deg > small_arcs) {
            next_run = LOOP_RUN_PUSH;    // a big level: scan + quad advance + bitmap absorb, in this same graph iteration
            break;
        }
    }

    // ---- hand the state back (CTA 0): to the kernels of this iteration, or to the host
    const bool aborted = ld_relaxed_gpu_u32(&sh->abort) != 0u;
    if (lead) {
        for (int i = 0; i < SMALL_SLOTS; ++i)
            for (int k = 0; k < B200_NUM_COUNTERS; ++k) a.pull_slots[i].c[k] = 0ull;
        s->level = level;
        s->m_unexplored = m_unexplored;
        s->flen = flen;
        s->reached = reached;
        s->total_arcs = total_arcs;
        s->bar_epoch = bar_epoch;
        s->stats_seq = stats_seq;
        s->launches = launches;
        s->pull = (next_run & LOOP_RUN_PULL) ? 1 : 0;
        s->dyn.in = in;
        s->dyn.out = out;
        s->dyn.len = len;
        s->dyn.bsel = bsel;
        s->dyn.next_label = level + 1;
        if (aborted) {
            s->timeout = 1u;             // the decide kernel of this iteration reports it and ends the loop
        } else if (finished) {
            a.res->status = status;
            a.res->num_levels = level;
            a.res->reached = reached;
            a.res->total_arcs = total_arcs;
            a.res->launches = launches;
            a.res->bar_epoch = bar_epoch;
            a.res->stats_seq = stats_seq;
            __threadfence_system();
            s->dyn.run = 0u;
            cudaGraphSetConditional(a.h_while, 0u);
        } else {
            s->dyn.run = next_run;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Big push level, receiver side: the exchange IS the `known` bitmap.  After the level's advance
// (BfsClaimPartQ: claim bits, emit nothing) and a flag barrier, every rank ORs ITS slice of every rank's
// `known` (peer loads over NVLink: (P-1)/P * n/8 bytes, whatever the frontier size), and what is new against
// `done` is labelled, appended to the next frontier (ascending ids) and written as the next frontier's
// bitmap slice -- one pass, word-wise, with the look-back scan of tile_scan.cuh.  Replaces the routed flush
// + publish + absorb + list -> slice of the vertex-id exchange, whose per-rank cost did not shrink with P.
// ---------------------------------------------------------------------------------------------
// DEG: also sum the degrees of the new vertices (the push -> pull decision of a direction-optimising traversal needs
// it; a push-only traversal does not, and the two dependent offset loads per vertex made this kernel latency-bound:
// 349 us for 14 M new vertices per rank in the 2-GPU trace)
template <int NT, int VT, bool DEG>
__global__ void __launch_bounds__(NT) p2p_absorb_bits_kernel(Peers peers, size_t off_known, int me, int P, uint32_t wl,
                                                             uint32_t *__restrict__ done, uint32_t *slice0, uint32_t *slice1,
                                                             int *__restrict__ labels, const uint32_t *__restrict__ offsets,
                                                             Partition part, P2PLoopState *s, SmallShared *sh,
                                                             unsigned long long capacity, unsigned long long *status,
                                                             unsigned long long *counters) {
    using TS = TileScan<NT, VT>;
    __shared__ typename TS::Smem sm;
    __shared__ uint32_t s_tile, s_bcast;
    const LoopDyn *dyn = &s->dyn;
    loop_trace(dyn, 5);
    if (!(dyn->run & LOOP_RUN_PUSH)) return;
    // ---- "my advance is done" to every peer; wait until theirs are (their claims are in their `known`)
    const unsigned long long epoch = s->bar_epoch + 1ull;
    if (P > 1) {
        if (blockIdx.x == 0 && threadIdx.x < (unsigned)P && (int)threadIdx.x != me) {
            __threadfence_system();
            st_release_sys(&reinterpret_cast<Ctrl *>(peers.base[threadIdx.x])->flags[me], epoch);
        }
        if (threadIdx.x < (unsigned)P && (int)threadIdx.x != me) {
            const unsigned long long *mine = &reinterpret_cast<const Ctrl *>(peers.base[me])->flags[threadIdx.x];
            unsigned long long t0 = 0ull;
            uint32_t polls = 0u;
            while (ld_acquire_sys(mine) < epoch) {
                if (spin_expired(polls, t0, P2P_TIMEOUT_NS)) {
                    atomicExch(&sh->abort, 1u);
                    break;
                }
                if ((polls & 63u) == 0u && ld_relaxed_gpu_u32(&sh->abort)) break;
            }
        }
        __syncthreads();
        if (ld_relaxed_gpu_u32(&sh->abort)) return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) loop_trace(dyn, 13);
    int *out = dyn->out;
    uint32_t *slice_w = dyn->bsel ? slice0 : slice1;
    const int next_label = dyn->next_label;
    uint32_t *known_own = reinterpret_cast<uint32_t *>(peers.base[me] + off_known) + (size_t)me * wl;
    unsigned int *tile_counter = &sh->absorb_tile;
    LookbackState st;
    st.status = status;
    st.tile_counter = tile_counter;
    st.epoch = (dyn->epoch + 1u) & 0x3FFFFFFFu;
    st.num_tiles = (wl + TS::NV - 1) / TS::NV;
    unsigned long long deg_sum = 0;
    // Level 0: the frontier was the source alone, so only its owner's `known` can hold new claims -- this rank reads its
    // slice of that ONE map (and its own) instead of all P (8 GPUs, scale 26: 2 MB instead of 8 MB of the 40 us this
    // kernel took at level 0, most of it over NVLink).
    const int only = s->level == 0 ? (int)part.owner((uint32_t)s->src) : -1;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= st.num_tiles) break;
        const uint32_t base = tile * TS::NV + (threadIdx.x >> 5) * (32 * VT) + lane_id();
        uint32_t word[VT], c[VT], ex[VT];
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            const uint32_t w = base + i * 32;
            uint32_t acc = 0u;
            if (w < wl) {
                if (only >= 0) {
                    acc = __ldcg(reinterpret_cast<const uint32_t *>(peers.base[only] + off_known) + (size_t)me * wl + w);
                    if (only != me) acc |= known_own[w];
                } else {
                    for (int q = 0; q < P; ++q)
                        acc |= __ldcg(reinterpret_cast<const uint32_t *>(peers.base[q] + off_known) + (size_t)me * wl + w);
                }
                const uint32_t d = done[w];
                word[i] = acc & ~d;
                if (word[i]) done[w] = d | word[i];
                known_own[w] = acc;
                slice_w[w] = word[i];
            } else {
                word[i] = 0u;
            }
            c[i] = __popc(word[i]);
        }
        const uint32_t total = TS::run(c, ex, sm);
        const uint32_t excl = lookback_exclusive(st, tile, total, &s_bcast);
        bool over = false;
#pragma unroll
        for (int i = 0; i < VT; ++i) {
            uint32_t bits = word[i];
            unsigned long long dest = (unsigned long long)excl + ex[i];
            const uint32_t r0 = (base + i * 32) << 5;
            while (bits) {
                const uint32_t b = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint32_t r = r0 + b;
                labels[r] = next_label;
                if (DEG) deg_sum += __ldg(offsets + r + 1) - __ldg(offsets + r);
                if (dest < capacity) out[dest] = (int)part.global_id((uint32_t)me, r);
                else over = true;
                ++dest;
            }
        }
        if (over) counters[B200_CNT_OVERFLOW] = 1ull;
        if (tile == st.num_tiles - 1 && threadIdx.x == 0) counters[B200_CNT_OUT] = (unsigned long long)excl + total;
        __syncthreads();
    }
#pragma unroll
    for (int d2 = 16; d2 > 0; d2 >>= 1) deg_sum += __shfl_xor_sync(FULL_MASK, deg_sum, d2);
    if (lane_id() == 0 && deg_sum) atomicAdd(&counters[B200_CNT_AUX], deg_sum);
}

__global__ void __launch_bounds__(256) p2p_gather_or_dyn_kernel(Peers peers, size_t off_slice0, size_t off_slice1,
                                                                uint32_t quads_per_slice, int P, int me, const LoopDyn *dyn,
                                                                uint32_t run_bit, uint4 *__restrict__ full,
                                                                uint4 *__restrict__ known, uint32_t *done,
                                                                const uint32_t *__restrict__ iso,
                                                                const uint32_t *__restrict__ pull_offsets) {
    loop_trace(dyn, run_bit == LOOP_RUN_PULL ? 7 : 10);
    if (!(dyn->run & run_bit)) return;
    if (run_bit == LOOP_RUN_PULL && (dyn->run & LOOP_RUN_TO_PULL)) {
        // first pull level after a big push level: rows without in-arcs count as done from here on (engine.cuh)
        or_no_in_arc_words(blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, pull_offsets, quads_per_slice * 128u, iso, done);
    }
    const size_t slice_off = dyn->bsel ? off_slice1 : off_slice0;
    const size_t total = (size_t)quads_per_slice * (size_t)P;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t p = (uint32_t)(i / quads_per_slice);
        const uint32_t j = (uint32_t)(i - (size_t)p * quads_per_slice);
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(peers.base[p] + slice_off) + j);
        full[i] = v;
        if (v.x | v.y | v.z | v.w) {
            uint4 k = known[i];
            k.x |= v.x; k.y |= v.y; k.z |= v.z; k.w |= v.w;
            known[i] = k;
        }
    }
    (void)me;
}

// ---------------------------------------------------------------------------------------------
// The pull phase of a direction-optimising traversal in ONE launch (the levels between the push -> pull switch and
// the hand-over back to push).  As kernels of the flat graph body a pull level cost ~40 us on top of its work at
// 2 GPUs (round-2 trace: four idle push nodes, the WHILE iteration, two idle hand-over nodes, a flag barrier with
// its kernel boundary, the decide kernel) -- three pull levels of a scale-26 traversal at 8 GPUs are 12 / 8 / 7 us
// of pull work each.  Per level here: gather + OR of every rank's frontier slice (peer loads: the allgather), grid
// barrier, early-exit pull over the local rows (bfs_pull_body, bitmaps read through L2), grid barrier, stats row +
// flag to every peer, wait, and the same decision in every CTA.  The level-closing flag barrier is the only
// cross-GPU synchronisation of a pull level: it also tells the peers that this rank's new slice is complete.
// ---------------------------------------------------------------------------------------------
constexpr int PULL_NT = 256;
#ifndef B200_P2P_PULL_CW
#define B200_P2P_PULL_CW 32
#endif
constexpr int PULL_CW = B200_P2P_PULL_CW;   // see bfs_pull_body
struct PullArgs {
    Peers peers;
    int me, P;
    Partition part;
    const uint32_t *pull_offsets;
    const int *pull_indices, *first_nbr;
    uint32_t *known, *full, *done;
    int *labels;
    size_t off_slice0, off_slice1;
    uint32_t wl;
    const uint32_t *iso;
    P2PLoopState *s;
    SmallShared *sh;
    PullCounters *slots;          // [SMALL_SLOTS], one per level (level & 3)
    unsigned int *tile_counters;  // workspace tile counters: [1] is re-armed for the hand-over compaction that follows
    P2PLoopResult *res;
    cudaGraphConditionalHandle h_while;
};

__global__ void __launch_bounds__(PULL_NT, 4) p2p_pull_levels_kernel(PullArgs a) {
    __shared__ unsigned s_gen;
    __shared__ unsigned long long s_sum[8];
    P2PLoopState *s = a.s;
    loop_trace(&s->dyn, 30);
    if (!(s->dyn.run & LOOP_RUN_PULL)) return;
    SmallShared *sh = a.sh;
    if (ld_relaxed_gpu_u32(&sh->abort)) return;
    const int me = a.me, P = a.P;
    const Partition part = a.part;
    const Ctrl *my_ctrl = reinterpret_cast<const Ctrl *>(a.peers.base[me]);
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    if (threadIdx.x == 0) s_gen = ld_relaxed_gpu_u32(&sh->bar_gen);
    uint32_t *slice[2] = {reinterpret_cast<uint32_t *>(a.peers.base[me] + a.off_slice0),
                          reinterpret_cast<uint32_t *>(a.peers.base[me] + a.off_slice1)};

    int level = s->level;
    const int mode = s->mode;
    const float beta = s->beta;
    const long long n = s->n, small_verts = s->small_verts;
    const uint32_t small_on = s->small_on;
    long long flen = s->flen, reached = s->reached, total_arcs = s->total_arcs, launches = s->launches;
    unsigned long long bar_epoch = s->bar_epoch;
    uint32_t stats_seq = s->stats_seq, bsel = s->dyn.bsel;
    int status = B200_OK;
    uint32_t next_run = 0u;
    unsigned long long next_local = 0ull;
    bool finished = false, first_level = true;
    // first pull level after a big push level: rows without in-arcs count as done from here on (engine.cuh); the grid
    // barrier after the gather orders this before the first read of `done`
    if (s->dyn.run & LOOP_RUN_TO_PULL)
        or_no_in_arc_words(blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, a.pull_offsets, part.n_local, a.iso, a.done);
    __syncthreads();

    for (;;) {
        PullCounters *c = &a.slots[level & (SMALL_SLOTS - 1)];
        if (lead) {
            PullCounters *z = &a.slots[(level + 1) & (SMALL_SLOTS - 1)];
            for (int i = 0; i < B200_NUM_COUNTERS; ++i) z->c[i] = 0ull;
        }
        // ---- the allgather: every rank's slice of the frontier -> full; folded into known
        loop_trace(&s->dyn, 31);
        {
            const size_t slice_off = bsel ? a.off_slice1 : a.off_slice0;
            const uint32_t qps = a.wl / 4;
            const size_t total = (size_t)qps * (size_t)P, stride = (size_t)gridDim.x * blockDim.x;
            uint4 *full4 = reinterpret_cast<uint4 *>(a.full), *known4 = reinterpret_cast<uint4 *>(a.known);
            // (four peer loads in flight per thread: one at a time is a chain of NVLink round trips -- 19-23 us per gather
            // of 7 MiB at 8 GPUs in the first trace)
            for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
                uint4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const size_t i = i0 + (size_t)k * stride;
                    v[k] = make_uint4(0u, 0u, 0u, 0u);
                    if (i < total) {
                        const uint32_t p = (uint32_t)(i / qps);
                        const uint32_t j = (uint32_t)(i - (size_t)p * qps);
                        v[k] = __ldcg(reinterpret_cast<const uint4 *>(a.peers.base[p] + slice_off) + j);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const size_t i = i0 + (size_t)k * stride;
                    if (i >= total) continue;
                    full4[i] = v[k];
                    if (v[k].x | v[k].y | v[k].z | v[k].w) {
                        uint4 kn = __ldcg(known4 + i);
                        kn.x |= v[k].x; kn.y |= v[k].y; kn.z |= v[k].z; kn.w |= v[k].w;
                        known4[i] = kn;
                    }
                }
            }
        }
        if (!small_grid_barrier(sh, &s_gen)) break;
        loop_trace(&s->dyn, 32);
        // The first pull level of a launch may read the bitmaps through L1 / the non-coherent path: nothing of `full` or `done`
        // was read earlier in this launch, so no stale line can exist (what the gather and the preset just wrote is in L2,
        // ordered by the grid barrier).  That level is the heavy one (28 M of 33 M discoveries at scale 26), and the L2-only
        // variant was 2x slower on it (444 vs ~210 us at scale 25 on one GPU).  Later levels of the launch must read through L2.
        if (first_level)
            bfs_pull_body<PULL_NT, false, PULL_CW>(part.n_local, a.pull_offsets, a.pull_indices, a.full, slice[bsel ^ 1u], a.done,
                                                   a.labels, level + 1, c->c, part, a.first_nbr);
        else
            bfs_pull_body<PULL_NT, true, PULL_CW>(part.n_local, a.pull_offsets, a.pull_indices, a.full, slice[bsel ^ 1u], a.done,
                                                  a.labels, level + 1, c->c, part, a.first_nbr);
        first_level = false;
        loop_trace(&s->dyn, 35);
        if (!small_grid_barrier(sh, &s_gen, true)) break;   // (system scope: the new slice, before the flag that says it is complete)
        loop_trace(&s->dyn, 36);
        // ---- level summary: row to every peer, flags, sums
        next_local = ld_volatile_u64(&c->c[B200_CNT_OUT]);
        const int parity = (int)(stats_seq & 1u);
        ++bar_epoch;
        if (blockIdx.x == 0 && threadIdx.x < (unsigned)P && (int)threadIdx.x != me) {
            Ctrl *pc = reinterpret_cast<Ctrl *>(a.peers.base[threadIdx.x]);
            unsigned long long *dst = pc->stats[parity][me];
            st_relaxed_sys(dst + ROW_NEXT, next_local);
            st_relaxed_sys(dst + ROW_ARCS, ld_volatile_u64(&c->c[B200_CNT_ARCS]));
            st_relaxed_sys(dst + ROW_DEG, ld_volatile_u64(&c->c[B200_CNT_AUX]));
            st_relaxed_sys(dst + ROW_SENT, 0ull);
            st_relaxed_sys(dst + ROW_OVERFLOW, 0ull);
            st_release_sys(&pc->flags[me], bar_epoch);
        }
        loop_trace(&s->dyn, 33);
        {   // wait for every peer's flag (local polling)
            if (threadIdx.x < (unsigned)P && (int)threadIdx.x != me) {
                const unsigned long long *mine = &my_ctrl->flags[threadIdx.x];
                unsigned long long t0 = 0ull;
                uint32_t polls = 0u;
                while (ld_acquire_sys(mine) < bar_epoch) {
                    if (spin_expired(polls, t0, P2P_TIMEOUT_NS)) {
                        atomicExch(&sh->abort, 1u);
                        break;
                    }
                    if ((polls & 63u) == 0u && ld_relaxed_gpu_u32(&sh->abort)) break;
                }
            }
            __syncthreads();
            if (ld_relaxed_gpu_u32(&sh->abort)) break;
        }
        loop_trace(&s->dyn, 34);
        if (threadIdx.x < 8) {
            unsigned long long t = 0;
            if (threadIdx.x == ROW_NEXT) t = next_local;
            else if (threadIdx.x == ROW_ARCS) t = ld_volatile_u64(&c->c[B200_CNT_ARCS]);
            else if (threadIdx.x == ROW_DEG) t = ld_volatile_u64(&c->c[B200_CNT_AUX]);
            for (int q = 0; q < P; ++q)
                if (q != me) t += ld_relaxed_sys(&my_ctrl->stats[parity][q][threadIdx.x]);
            s_sum[threadIdx.x] = t;
        }
        __syncthreads();
        const long long found = (long long)s_sum[ROW_NEXT], arcs = (long long)s_sum[ROW_ARCS];
        __syncthreads();
        if (lead && level < B200_MAX_LEVELS) {
            P2PLevelRec *r = &a.res->level[level];
            r->direction = 1;
            r->exchange = 1;
            r->frontier_len = flen;
            r->arcs = arcs;
            r->discovered = found;
            r->sent = 0;
        }
        loop_trace(&s->dyn, 64u + ((unsigned)level & 63u));   // level `level` closed
        total_arcs += arcs;
        launches += 1;
        ++level;
        ++stats_seq;
        if (found == 0) {
            finished = true;
            break;
        }
        reached += found;
        bsel ^= 1u;
        const bool hand_over = mode == B200_BFS_BEAMER && (double)found < (double)n / beta && found < flen;
        flen = found;
        if (hand_over) {
            // Back to push, inside this launch: this rank's slice of the new frontier becomes its frontier list (order is
            // free: warp-aggregated appends, no scan) and is folded into known[own slice] -- the push kernels rely on
            // known >= done for owned vertices.  The peers' slices are NOT folded: a vertex they found in this last pull
            // level may be claimed and offered to its owner once more, and the owner's `done` turns it away.
            const unsigned lane = lane_id();
            const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, total_warps = (gridDim.x * blockDim.x) >> 5;
            const uint32_t *fr = slice[bsel];
            uint32_t *known_own = a.known + (size_t)me * a.wl;
            int *list = const_cast<int *>(s->dyn.in);
            unsigned long long *cnt = &c->c[B200_CNT_AUX2];
            for (uint32_t w0 = gwarp * 32u; w0 < a.wl; w0 += total_warps * 32u) {
                const uint32_t w = w0 + lane;
                uint32_t bits = w < a.wl ? __ldcg(fr + w) : 0u;
                if (bits) known_own[w] = __ldcg(known_own + w) | bits;
                const uint32_t k = __popc(bits);
                uint32_t incl = k;
#pragma unroll
                for (int d2 = 1; d2 < 32; d2 <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL_MASK, incl, d2);
                    if (lane >= (unsigned)d2) incl += t;
                }
                const uint32_t tot = __shfl_sync(FULL_MASK, incl, 31);
                if (tot == 0u) continue;
                unsigned long long base = 0;
                if (lane == 31) base = atomicAdd(cnt, (unsigned long long)tot);
                base = __shfl_sync(FULL_MASK, base, 31) + incl - k;
                while (bits) {
                    const uint32_t b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (base < part.n_local) list[base] = (int)part.global_id((uint32_t)me, (w << 5) + b);
                    ++base;
                }
            }
            if (!small_grid_barrier(sh, &s_gen)) break;
            next_run = (small_on && found <= small_verts) ? LOOP_RUN_SMALL : LOOP_RUN_PUSH;
            break;
        }
    }

    const bool aborted = ld_relaxed_gpu_u32(&sh->abort) != 0u;
    if (lead) {
        for (int i = 0; i < SMALL_SLOTS; ++i) small_zero_slot(&sh->slot[i]);   // the small-level kernel may run next
        a.tile_counters[1] = 0u;
        s->level = level;
        s->flen = flen;
        s->reached = reached;
        s->total_arcs = total_arcs;
        s->bar_epoch = bar_epoch;
        s->stats_seq = stats_seq;
        s->launches = launches;
        s->pull = 0;
        s->dyn.bsel = bsel;
        s->dyn.next_label = level + 1;
        s->dyn.len = (uint32_t)next_local;
        s->dyn.epoch = (s->dyn.epoch + 2u) & 0x3FFFFFFFu;   // the hand-over compaction gets a look-back tag nobody used
        if (aborted) {
            s->timeout = 1u;             // the decide kernel of the next iteration reports it and ends the loop
        } else if (finished) {
            a.res->status = status;
            a.res->num_levels = level;
            a.res->reached = reached;
            a.res->total_arcs = total_arcs;
            a.res->launches = launches;
            a.res->bar_epoch = bar_epoch;
            a.res->stats_seq = stats_seq;
            __threadfence_system();
            s->dyn.run = 0u;
            cudaGraphSetConditional(a.h_while, 0u);
        } else {
            s->dyn.run = next_run;
        }
    }
}

// Level summary of a big push / pull level (this rank's row into every heap, flag barrier, sum of the rows = the
// allreduce) + the host loop's decision, on every rank identically.
__global__ void __launch_bounds__(32) p2p_stats_decide_kernel(Peers peers, int me, int P, P2PLoopState *s,
                                                              unsigned long long *counters, unsigned int *tile_counters,
                                                              SmallShared *sh, PullCounters *pull_slots, P2PLoopResult *res,
                                                              cudaGraphConditionalHandle h_while) {
    const int p = (int)lane_id();
    loop_trace(&s->dyn, 9);
    const uint32_t run = s->dyn.run;
    const bool aborted = ld_relaxed_gpu_u32(&sh->abort) != 0u || s->timeout != 0u;
    if (!(run & LOOP_RUN_PUSH) || aborted) {
        // the small-level and the pull-levels kernels close their levels themselves; only a time-out is left to report
        if (aborted && p == 0) {
            res->status = B200_ERR_TIMEOUT;
            res->num_levels = s->level;
            res->reached = s->reached;
            res->total_arcs = s->total_arcs;
            res->launches = s->launches;
            res->bar_epoch = s->bar_epoch;
            res->stats_seq = s->stats_seq;
            __threadfence_system();
            s->dyn.run = 0u;
            cudaGraphSetConditional(h_while, 0u);
        }
        return;
    }
    const bool was_pull = false;   // (pull levels: p2p_pull_levels_kernel)
    // a big push level used one epoch for its "advance done" barrier (p2p_absorb_bits_kernel)
    const unsigned long long epoch = s->bar_epoch + ((!was_pull && P > 1) ? 2ull : 1ull);
    const int parity = (int)(s->stats_seq & 1u);
    __syncwarp();
    unsigned long long row[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) row[k] = 0;
    row[ROW_NEXT] = counters[B200_CNT_OUT];
    row[ROW_ARCS] = counters[B200_CNT_ARCS];
    row[ROW_DEG] = counters[B200_CNT_AUX];
    row[ROW_OVERFLOW] = counters[B200_CNT_OVERFLOW];
    if (p < P) {
        unsigned long long *dst = reinterpret_cast<Ctrl *>(peers.base[p])->stats[parity][me];
#pragma unroll
        for (int k = 0; k < 8; ++k) st_relaxed_sys(dst + k, row[k]);
    }
    const bool ok = p2p_signal_wait(peers, me, P, epoch);
    if (p == 0) loop_trace(&s->dyn, 12);
    unsigned long long mysum = 0;
    if (p < 8) {
        const Ctrl *mine = reinterpret_cast<const Ctrl *>(peers.base[me]);
        for (int q = 0; q < P; ++q) mysum += ld_relaxed_sys(&mine->stats[parity][q][p]);
    }
    const long long found = (long long)__shfl_sync(FULL_MASK, mysum, ROW_NEXT);
    const long long arcs = (long long)__shfl_sync(FULL_MASK, mysum, ROW_ARCS);
    const long long next_deg = (long long)__shfl_sync(FULL_MASK, mysum, ROW_DEG);
    const bool overflow = __shfl_sync(FULL_MASK, mysum, ROW_OVERFLOW) != 0ull;
    // re-arm the small-level kernel's counter slots (it may run in the next iteration)
    for (int i = p; i < SMALL_SLOTS; i += 32) {
        small_zero_slot(&sh->slot[i]);
        for (int k = 0; k < B200_NUM_COUNTERS; ++k) pull_slots[i].c[k] = 0ull;
    }
    if (p != 0) return;

    const long long next_local = (long long)row[ROW_NEXT];
    bool pull = was_pull;
    int level = s->level;
    if (level < B200_MAX_LEVELS) {
        P2PLevelRec *r = &res->level[level];
        r->direction = was_pull ? 1 : 0;
        r->exchange = 1;
        r->frontier_len = s->flen;
        r->arcs = arcs;
        r->discovered = found;
        r->sent = 0;
    }
    loop_trace(&s->dyn, 64u + ((unsigned)level & 63u));   // level `level` closed
    s->total_arcs += arcs;
    s->launches += s->kernels_per_level;
    ++level;
    s->level = level;
    s->bar_epoch = epoch;
    s->stats_seq += 1u;
    bool done = false;
    uint32_t trans = 0u;
    int status = B200_OK;
    bool next_small = false;
    if (!ok) {
        status = B200_ERR_TIMEOUT;
        done = true;
    } else if (overflow) {
        status = B200_ERR_OVERFLOW;
        done = true;
    } else if (found == 0) {
        done = true;
    } else {
        const long long flen = s->flen;
        s->reached += found;
        if (!was_pull) {
            s->m_unexplored -= arcs;
            const int *t = s->dyn.in;
            s->dyn.in = s->dyn.out;
            s->dyn.out = const_cast<int *>(t);
            s->dyn.len = (uint32_t)next_local;
            if (s->mode == B200_BFS_BEAMER && (double)next_deg > (double)s->m_unexplored / s->alpha && found > flen) {
                pull = true;
                trans = LOOP_RUN_TO_PULL;   // (the gather of the first pull level also applies the no-in-arc preset)
                s->dyn.bsel ^= 1u;          // the slice written this level is the frontier of the first pull level
            } else {
                // (push-only traversals do not sum the degrees of what they find: the vertex count stands in)
                next_small = s->small_on && (s->mode == B200_BFS_BEAMER ? next_deg <= s->small_arcs : found <= s->small_verts);
            }
        } else {
            s->dyn.bsel ^= 1u;
            if ((double)found < (double)s->n / s->beta && found < flen) {
                // hand the (small) frontier back to push: the last pull discoveries of the peers become
                // "known", this rank's slice becomes its frontier list (gather + compaction below)
                trans = LOOP_RUN_TO_PUSH;
                s->dyn.len = (uint32_t)next_local;
                pull = false;
                next_small = s->small_on && found <= s->small_verts;   // (pull levels do not sum the degrees of what they find)
            }
        }
        s->flen = found;
    }
    s->pull = pull ? 1 : 0;
    s->dyn.next_label = level + 1;
    s->dyn.epoch = (s->dyn.epoch + 2u) & 0x3FFFFFFFu;
    for (int i = 0; i < B200_NUM_COUNTERS; ++i) counters[i] = 0ull;
    tile_counters[0] = 0u;
    tile_counters[1] = 0u;
    sh->absorb_tile = 0u;
    if (done) {
        res->status = status;
        res->num_levels = level;
        res->reached = s->reached;
        res->total_arcs = s->total_arcs;
        res->launches = s->launches;
        res->bar_epoch = s->bar_epoch;
        res->stats_seq = s->stats_seq;
        __threadfence_system();
    }
    s->dyn.run = done ? 0u : ((pull ? LOOP_RUN_PULL : (next_small ? LOOP_RUN_SMALL : LOOP_RUN_PUSH)) | trans);
    cudaGraphSetConditional(h_while, done ? 0u : 1u);
}

// CUDA loads kernels lazily, and a load may have to wait for every kernel running in the context.
// A rank that spins in a flag barrier while a peer rank OF THE SAME PROCESS (ranks as threads) is
// about to launch a kernel for the first time would therefore deadlock: load everything up front.
template <class K>
cudaError_t preload(K k) {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, k);
}
cudaError_t preload_kernels() {
    cudaError_t e;
    constexpr int VT = B200_QUAD_VT;
    if ((e = preload_no_in_arc_kernel()) != cudaSuccess) return e;
    if ((e = preload(p2p_init_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_publish_counts_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_stats_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_absorb_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_list_to_slice_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_gather_or_kernel)) != cudaSuccess) return e;
    if ((e = preload(bfs_pull_kernel<256>)) != cudaSuccess) return e;
    if ((e = preload(scan_sizes_kernel<SCAN_NT, SCAN_VT, FrontierQuads>)) != cudaSuccess) return e;
    if ((e = preload(scan_sizes_kernel<SCAN_NT, SCAN_VT, FrontierDegree>)) != cudaSuccess) return e;
    if ((e = preload(bitmap_list_kernel<COMPACT_NT, BITLIST_VT, BitmapWords, PartItem>)) != cudaSuccess) return e;
    if ((e = preload(quad_advance_kernel<BfsPushPartQ, OUT_ROUTED, true, QUAD_NT, VT, QUAD_WSEG>)) != cudaSuccess) return e;
    if ((e = preload(scan_sizes_dyn_kernel<SCAN_NT, SCAN_VT, FrontierQuadsDynPart>)) != cudaSuccess) return e;
    if ((e = preload(bitmap_list_dyn_kernel<COMPACT_NT, BITLIST_VT, BitmapWordsDyn, PartItem>)) != cudaSuccess) return e;
    if ((e = preload(p2p_loop_init_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_small_levels_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_pull_levels_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_absorb_bits_kernel<COMPACT_NT, BITLIST_VT, true>)) != cudaSuccess) return e;
    if ((e = preload(p2p_absorb_bits_kernel<COMPACT_NT, BITLIST_VT, false>)) != cudaSuccess) return e;
    if ((e = preload(quad_advance_kernel<BfsClaimPartQ, OUT_NONE, false, QUAD_NT, VT, QUAD_WSEG>)) != cudaSuccess) return e;
    if ((e = preload(p2p_gather_or_dyn_kernel)) != cudaSuccess) return e;
    if ((e = preload(p2p_stats_decide_kernel)) != cudaSuccess) return e;
    if ((e = preload(lbs_advance_kernel<BfsPushPartOp, OUT_ROUTED, true, LBS_NT, LBS_VT, LBS_SEG_T>)) != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace

struct b200_p2p_bfs {
    b200_ctx *ctx;
    int rank, P;
    uint32_t log_p;
    int64_t n_global, n_local;
    uint32_t wl;                       // words per bitmap slice
    unsigned char *heap;
    size_t heap_bytes, off_slice[2], off_known, off_inbox;
    bool connected;
    Peers peers;
    void *opened[P2P_MAX];             // cudaIpcOpenMemHandle results (closed in destroy)
    uint32_t *known, *full;            // [n_global/32] each; known lives in the heap (peers read their slices of it), full is local
    uint32_t *done;                    // [n_local/32] labelled set of the owned vertices (graph-driven loop)
    SmallShared *small;                // counters + grid barrier of the small-level kernel
    uint32_t *big_rows;                // [n_local] long rows of a small level
    int small_grid;                    // CTAs of the small-level kernel (all resident)
    PullCounters *pull_slots;          // [SMALL_SLOTS] per-level counters of the pull-levels kernel
    int pull_grid;                     // CTAs of the pull-levels kernel (all resident)
    unsigned long long *absorb_status; // look-back tile states of p2p_absorb_bits_kernel (its own array: the scan and the
    int64_t absorb_tiles;              // hand-over compaction share the workspace's, with tags of their own)
    unsigned long long *box_counts;    // [P2P_MAX] device counters of the routed boxes
    unsigned long long *h_res, *d_res; // mapped pinned result block
    unsigned long long epoch;
    unsigned stats_seq;
    cudaEvent_t ev_run[2];
    cudaEvent_t ev_level[P2P_TIMED_LEVELS + 1];
    bool have_level_events;
    // graph-driven level loop
    P2PLoopState *d_lstate;
    P2PLoopParams *h_lparams, *d_lparams;
    P2PLoopResult *h_lresult, *d_lresult;
    cudaStream_t cap_stream;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    unsigned long long *d_trace;
    unsigned long long *h_trace;       // host copy of the last traced run: [0] = entries, then (globaltimer ns << 8 | id)
    int trace_on;                      // b200_p2p_bfs_set_trace
    const void *k_offsets, *k_indices, *k_labels, *k_scratch, *k_iso, *k_first, *k_pull_offsets, *k_pull_indices;
    uint64_t k_gen;            // b200_ctx::scratch_gen the graph was built against
    int k_mode;
    int graph_failed;
};

namespace {

void p2p_drop_graph(b200_p2p_bfs *s) {
    if (s->exec) cudaGraphExecDestroy(s->exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    s->exec = nullptr;
    s->graph = nullptr;
    s->k_offsets = s->k_indices = s->k_labels = s->k_scratch = s->k_iso = s->k_first = nullptr;
    s->k_pull_offsets = s->k_pull_indices = nullptr;
}

int p2p_build_graph(b200_p2p_bfs *s, const b200_graph *g, int32_t *d_labels, int mode) {
    b200_ctx *ctx = s->ctx;
    b200_workspace *ws = &ctx->ws;
    const int me = s->rank, P = s->P;
    const Partition part{s->log_p, (uint32_t)me, (uint32_t)s->n_local};
    const bool beamer = mode == B200_BFS_BEAMER;
    const uint32_t *pull_off = g->col_offsets ? g->col_offsets : g->row_offsets;
    const int32_t *pull_idx = g->row_indices ? g->row_indices : g->col_indices;
    uint32_t *slice0 = reinterpret_cast<uint32_t *>(s->heap + s->off_slice[0]);
    uint32_t *slice1 = reinterpret_cast<uint32_t *>(s->heap + s->off_slice[1]);
    cudaStream_t cs = s->cap_stream;
    void *const user_stream = ws->stream;
    const int64_t launches0 = ws->launches;
    int st = B200_OK;
    bool capturing = false;
    cudaGraph_t G = nullptr, body = nullptr;
    cudaGraphConditionalHandle h_while;
    cudaGraphNode_t deps[8], n_while;
    size_t ndeps = 0;
    cudaGraphNodeParams np_w = {};
    const LoopDyn *dyn = &s->d_lstate->dyn;
    const int kernels_per_level = 4 + (beamer ? 4 : 0);

    p2p_drop_graph(s);
    ws->stream = (void *)cs;
    LL_CUDA(cudaGraphCreate(&G, 0));
    LL_CUDA(cudaGraphConditionalHandleCreate(&h_while, G, 1, cudaGraphCondAssignDefault));

    // the persistent small-level kernel (consecutive SMALL push levels, vertex ids through the inboxes): once in the
    // prologue (level 0 is always small) and as the LAST node of the loop body, so that the iteration it finishes the
    // traversal in ends there -- no trailing idle nodes
    SmallArgs sa;
    std::memset(&sa, 0, sizeof sa);
    sa.peers = s->peers;
    sa.me = me;
    sa.P = P;
    sa.part = part;
    sa.offsets = g->row_offsets;
    sa.indices = g->col_indices;
    sa.known = s->known;
    sa.done = s->done;
    sa.labels = d_labels;
    sa.s = s->d_lstate;
    sa.sh = s->small;
    sa.big = s->big_rows;
    sa.off_inbox = s->off_inbox;
    sa.off_slice0 = s->off_slice[0];
    sa.off_slice1 = s->off_slice[1];
    sa.wl = s->wl;
    sa.iso = g->no_in_arc_bitmap;
    sa.pull_offsets = pull_off;
    sa.res = s->d_lresult;
    sa.h_while = h_while;
    sa.pull_slots = s->pull_slots;

    // ---- prologue
    LL_CUDA(cudaStreamBeginCaptureToGraph(cs, G, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    capturing = true;
    LL_CUDA(cudaMemsetAsync(d_labels, 0xFF, sizeof(int32_t) * (size_t)s->n_local, cs));
    LL_CUDA(cudaMemsetAsync(s->known, 0, (size_t)(s->n_global / 8), cs));
    LL_CUDA(cudaMemsetAsync(s->done, 0, sizeof(uint32_t) * (size_t)s->wl, cs));
    LL_CUDA(cudaMemsetAsync(s->heap + s->off_slice[0], 0, s->off_known - s->off_slice[0], cs));   // both frontier slices
    p2p_loop_init_kernel<<<1, 1, 0, cs>>>(s->d_lparams, s->d_lstate, d_labels, s->known, s->done, ctx->frontier[0], ctx->frontier[1],
                                          (long long)s->n_global, part, ws->d_counters, ws->d_tile_counter, s->small,
                                          s->pull_slots, kernels_per_level);
    LL_CUDA(cudaGetLastError());
    p2p_small_levels_kernel<<<(unsigned)s->small_grid, SMALL_NT, SMALL_SMEM, cs>>>(sa);
    LL_CUDA(cudaGetLastError());
    capturing = false;
    if ((st = end_capture(cs, deps, &ndeps, 8)) != B200_OK) goto fail;

    np_w.type = cudaGraphNodeTypeConditional;
    np_w.conditional.handle = h_while;
    np_w.conditional.type = cudaGraphCondTypeWhile;
    np_w.conditional.size = 1;
    LL_CUDA(cudaGraphAddNode(&n_while, G, deps, ndeps, &np_w));
    body = np_w.conditional.phGraph_out[0];

    // ---- the flat body of one graph iteration: every kernel returns at once unless its LOOP_RUN_* bit is set
    LL_CUDA(cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    capturing = true;
    {
        // (2) a BIG push level: quad scan, quad advance that only claims bits in `known`, then the bitmap exchange
        const int64_t max_tiles = (s->n_local + SCAN_NT * SCAN_VT - 1) / (SCAN_NT * SCAN_VT);
        int64_t grid = (int64_t)ws->num_sms * 8;
        if (grid > max_tiles) grid = max_tiles;
        FrontierQuadsDynPart fn{dyn, g->row_offsets, reinterpret_cast<uint2 *>(ws->d_rows), part.log_p};
        scan_sizes_dyn_kernel<SCAN_NT, SCAN_VT><<<(unsigned)grid, SCAN_NT, 0, cs>>>(
            fn, dyn, (uint32_t)LOOP_RUN_PUSH, ws->d_scanned, ws->d_status, ws->d_tile_counter, ws->d_counters + B200_CNT_TOTAL);
        LL_CUDA(cudaGetLastError());
        QuadArgs a = make_quad_args(ws, ctx->frontier[0], 0u, g->row_offsets, g->col_indices, nullptr, part.log_p);
        a.dyn = dyn;
        BfsClaimPartQ op{s->known, part};
        LL_CUDA((launch_quad_advance<OUT_NONE, false>(ws, a, op, nullptr, 0ull)));
        const int64_t atiles = ((int64_t)s->wl + COMPACT_NT * BITLIST_VT - 1) / (COMPACT_NT * BITLIST_VT);
        int64_t agrid = (int64_t)ws->num_sms * 4;   // (a persistent tile loop; an idle node of 148 x 8 CTAs cost 3.7 us)
        if (agrid > atiles) agrid = atiles;
        if (beamer)
            p2p_absorb_bits_kernel<COMPACT_NT, BITLIST_VT, true><<<(unsigned)agrid, COMPACT_NT, 0, cs>>>(
                s->peers, s->off_known, me, P, s->wl, s->done, slice0, slice1, d_labels, g->row_offsets, part, s->d_lstate, s->small,
                (unsigned long long)s->n_local, s->absorb_status, ws->d_counters);
        else
            p2p_absorb_bits_kernel<COMPACT_NT, BITLIST_VT, false><<<(unsigned)agrid, COMPACT_NT, 0, cs>>>(
                s->peers, s->off_known, me, P, s->wl, s->done, slice0, slice1, d_labels, g->row_offsets, part, s->d_lstate, s->small,
                (unsigned long long)s->n_local, s->absorb_status, ws->d_counters);
        LL_CUDA(cudaGetLastError());
        // (3) level summary across the ranks + decision of a big push level
        p2p_stats_decide_kernel<<<1, 32, 0, cs>>>(s->peers, me, P, s->d_lstate, ws->d_counters, ws->d_tile_counter, s->small,
                                                  s->pull_slots, s->d_lresult, h_while);
        LL_CUDA(cudaGetLastError());
        if (beamer) {
            // (4) the whole pull phase in one persistent kernel (after the push -> pull switch of (1) or (3))
            PullArgs pa;
            std::memset(&pa, 0, sizeof pa);
            pa.peers = s->peers;
            pa.me = me;
            pa.P = P;
            pa.part = part;
            pa.pull_offsets = pull_off;
            pa.pull_indices = pull_idx;
            pa.first_nbr = g->first_in_neighbor;
            pa.known = s->known;
            pa.full = s->full;
            pa.done = s->done;
            pa.labels = d_labels;
            pa.off_slice0 = s->off_slice[0];
            pa.off_slice1 = s->off_slice[1];
            pa.wl = s->wl;
            pa.iso = g->no_in_arc_bitmap;
            pa.s = s->d_lstate;
            pa.sh = s->small;
            pa.slots = s->pull_slots;
            pa.tile_counters = ws->d_tile_counter;
            pa.res = s->d_lresult;
            pa.h_while = h_while;
            p2p_pull_levels_kernel<<<(unsigned)s->pull_grid, PULL_NT, 0, cs>>>(pa);
            LL_CUDA(cudaGetLastError());
        }
        // (5) the small levels that follow (the tail of the traversal, usually)
        p2p_small_levels_kernel<<<(unsigned)s->small_grid, SMALL_NT, SMALL_SMEM, cs>>>(sa);
        LL_CUDA(cudaGetLastError());
    }
    capturing = false;
    if ((st = end_capture(cs, deps, &ndeps, 8)) != B200_OK) goto fail;

    LL_CUDA(cudaGraphInstantiate(&s->exec, G, 0));
    s->graph = G;
    s->k_offsets = g->row_offsets;
    s->k_indices = g->col_indices;
    s->k_labels = d_labels;
    s->k_scratch = ctx->frontier[0];
    s->k_iso = g->no_in_arc_bitmap;
    s->k_first = g->first_in_neighbor;
    s->k_pull_offsets = pull_off;
    s->k_pull_indices = pull_idx;
    s->k_gen = ctx->scratch_gen;
    s->k_mode = mode;
    ws->stream = user_stream;
    ws->launches = launches0;
    return B200_OK;

fail:
    if (capturing) {
        cudaGraph_t junk = nullptr;
        cudaStreamEndCapture(cs, &junk);
    }
    (void)cudaGetLastError();
    if (G) cudaGraphDestroy(G);
    ws->stream = user_stream;
    ws->launches = launches0;
    return st;
}

// Builds (or reuses) the graph for (g, labels, mode).  B200_ERR_UNSUPPORTED if it cannot be built here.
int p2p_ensure_graph(b200_p2p_bfs *s, const b200_graph *g, int32_t *d_labels, int mode) {
    b200_ctx *ctx = s->ctx;
    cudaStream_t st = ws_stream(&ctx->ws);
    if (s->graph_failed) return B200_ERR_UNSUPPORTED;
    if (!s->cap_stream) {
        B200_CUDA(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
        B200_CUDA(cudaMalloc(&s->d_lstate, sizeof(P2PLoopState)));
        B200_CUDA(cudaHostAlloc(&s->h_lparams, sizeof(P2PLoopParams), cudaHostAllocMapped));
        B200_CUDA(cudaHostAlloc(&s->h_lresult, sizeof(P2PLoopResult), cudaHostAllocMapped));
        B200_CUDA(cudaHostGetDevicePointer(&s->d_lparams, s->h_lparams, 0));
        B200_CUDA(cudaHostGetDevicePointer(&s->d_lresult, s->h_lresult, 0));
    }
    if (!s->exec || s->k_offsets != g->row_offsets || s->k_indices != g->col_indices || s->k_labels != d_labels ||
        s->k_scratch != ctx->frontier[0] || s->k_mode != mode || s->k_iso != g->no_in_arc_bitmap ||
        s->k_first != g->first_in_neighbor || s->k_gen != ctx->scratch_gen ||
        s->k_pull_offsets != (g->col_offsets ? g->col_offsets : g->row_offsets) ||
        s->k_pull_indices != (g->row_indices ? g->row_indices : g->col_indices)) {
        B200_CUDA(cudaStreamSynchronize(st));
        const int bs = p2p_build_graph(s, g, d_labels, mode);
        if (bs != B200_OK) {
            s->graph_failed = bs;
            return B200_ERR_UNSUPPORTED;
        }
        B200_CUDA(cudaGraphUpload(s->exec, st));
        B200_CUDA(cudaStreamSynchronize(st));
    }
    return B200_OK;
}

// b200_p2p_bfs_run through the graph; B200_ERR_UNSUPPORTED => the caller runs the host-driven loop.
int p2p_run_graph(b200_p2p_bfs *s, const b200_graph *g, int64_t m_global, int32_t src, int mode, float alpha, float beta,
                  int32_t *d_labels, b200_stats *stats, int64_t *sent_per_level) {
    b200_ctx *ctx = s->ctx;
    b200_workspace *ws = &ctx->ws;
    cudaStream_t st = ws_stream(ws);
    B200_TRY(p2p_ensure_graph(s, g, d_labels, mode));
    if (ws->epoch > 0x3FFFFFFFu - 4u * (B200_MAX_LEVELS + 2)) {
        B200_CUDA(cudaMemsetAsync(ws->d_status, 0, sizeof(unsigned long long) * (size_t)ws->status_tiles, st));
        ws->epoch = 0;
    }
    P2PLoopParams *pr = s->h_lparams;
    pr->src = src;
    pr->mode = mode;
    pr->alpha = alpha;
    pr->beta = beta;
    pr->m_global = m_global;
    pr->bar_epoch0 = s->epoch;
    pr->lb_epoch0 = ws->epoch + 1;
    pr->stats_seq0 = s->stats_seq;
    const bool want_trace = getenv("B200_LOOP_TRACE") != nullptr || s->trace_on;   // (read per run)
    constexpr uint32_t TRACE_CAP = 4096;
    if (want_trace && !s->d_trace) B200_CUDA(cudaMalloc(&s->d_trace, sizeof(unsigned long long) * TRACE_CAP));
    pr->trace = want_trace ? s->d_trace : nullptr;
    pr->trace_cap = TRACE_CAP;
    // small-level kernel: on unless B200_P2P_SMALL=0; a push level is "small" while the degree sum of its frontier
    // stays below B200_P2P_SMALL_ARCS (after a pull phase: while the frontier has at most B200_P2P_SMALL_VERTS vertices)
    // (read per run: all ranks of a job see the same environment)
    const long long env_small = getenv("B200_P2P_SMALL") ? atoll(getenv("B200_P2P_SMALL")) : 1;
    const long long env_arcs = getenv("B200_P2P_SMALL_ARCS") ? atoll(getenv("B200_P2P_SMALL_ARCS")) : (2ll << 20);
    const long long env_verts = getenv("B200_P2P_SMALL_VERTS") ? atoll(getenv("B200_P2P_SMALL_VERTS")) : (64ll << 10);
    pr->hub_min = getenv("B200_P2P_HUB_MIN") ? atoll(getenv("B200_P2P_HUB_MIN")) : (32ll << 10);
    pr->small_on = env_small != 0 ? 1u : 0u;
    pr->small_arcs = env_arcs;
    pr->small_verts = env_verts;
    s->h_lresult->status = -1;
    s->h_lresult->num_levels = 0;
    B200_CUDA(cudaEventRecord(s->ev_run[0], st));
    B200_CUDA(cudaGraphLaunch(s->exec, st));
    B200_CUDA(cudaEventRecord(s->ev_run[1], st));
    B200_CUDA(cudaEventSynchronize(s->ev_run[1]));
    const P2PLoopResult *r = s->h_lresult;
    if (r->status < 0) return B200_ERR_CUDA;
    const char *tenv = getenv("B200_LOOP_TRACE");
    const bool trace_all = tenv && !strcmp(tenv, "all");
    if (want_trace) {
        if (!s->h_trace) s->h_trace = new (std::nothrow) unsigned long long[TRACE_CAP];
        if (!s->h_trace) return B200_ERR_NOMEM;
        B200_CUDA(cudaMemcpy(s->h_trace, s->d_trace, sizeof(unsigned long long) * TRACE_CAP, cudaMemcpyDeviceToHost));
        if (s->h_trace[0] > TRACE_CAP - 1) s->h_trace[0] = TRACE_CAP - 1;
    }
    if (tenv && (s->rank == 0 || trace_all)) {   // debug aid: kernel-entry timeline (ids: see loop_trace call sites)
        const unsigned long long *h = s->h_trace;
        const unsigned long long cnt = h[0];
        fprintf(stderr, "B200_LOOP_TRACE rank%d %llu entries (us since first, kernel id):", s->rank, cnt);
        for (unsigned long long i = 1; i <= cnt; ++i)
            fprintf(stderr, " %.1f:%llu", (double)((h[i] >> 8) - (h[1] >> 8)) * 1e-3, h[i] & 255ull);
        fprintf(stderr, "\n");
    }
    s->epoch = r->bar_epoch;
    s->stats_seq = r->stats_seq;
    const int levels = r->num_levels;
    const unsigned used = 2u * (unsigned)levels + 2u;
    if (used > 0x3FFFFFFFu - ws->epoch - 2u) {
        B200_CUDA(cudaMemsetAsync(ws->d_status, 0, sizeof(unsigned long long) * (size_t)ws->status_tiles, st));
        ws->epoch = 0;
    } else {
        ws->epoch += used;
    }
    ws->launches += r->launches;
    const int nl = levels < B200_MAX_LEVELS ? levels : B200_MAX_LEVELS;
    if (stats) {
        stats->num_levels = levels;
        stats->reached = r->reached;
        stats->total_arcs = r->total_arcs;
        stats->launches = r->launches;
        stats->level_loop = B200_LOOP_GRAPH;
        B200_CUDA(cudaEventElapsedTime(&stats->device_ms, s->ev_run[0], s->ev_run[1]));
        for (int l = 0; l < nl; ++l) {
            b200_level_stat *ls = &stats->level[l];
            ls->direction = r->level[l].direction;
            ls->reserved = r->level[l].exchange;   // 0 = vertex ids through the inboxes (small level), 1 = bitmap slices
            ls->frontier_len = r->level[l].frontier_len;
            ls->arcs = r->level[l].arcs;
            ls->discovered = r->level[l].discovered;
            ls->advance_ms = 0.f;
            ls->level_ms = 0.f;
        }
    }
    if (sent_per_level)
        for (int l = 0; l < nl; ++l) sent_per_level[l] = r->level[l].sent;
    return r->status;
}

}  // namespace

extern "C" {

int b200_p2p_bfs_heap_bytes(int num_ranks, int64_t n_global, int64_t *bytes) {
    if (!bytes || num_ranks < 1 || num_ranks > P2P_MAX || (num_ranks & (num_ranks - 1)) || n_global < 1) return B200_ERR_INVALID;
    const int64_t n_local = n_global / num_ranks;
    if (n_local * num_ranks != n_global || n_local % 128) return B200_ERR_INVALID;
    const size_t slice = ((size_t)(n_local / 32) * 4 + 255) & ~(size_t)255;
    const size_t known = ((size_t)(n_global / 8) + 255) & ~(size_t)255;
    *bytes = (int64_t)(P2P_CTRL_BYTES + 2 * slice + known + (size_t)n_global * 4);
    return B200_OK;
}

int b200_p2p_bfs_create(b200_ctx *ctx, int rank, int num_ranks, int64_t n_global, b200_p2p_bfs **out,
                        void *ipc_handle_out, void **heap_base_out) {
    if (!ctx || !out || rank < 0 || rank >= num_ranks || n_global > (1ll << 31)) return B200_ERR_INVALID;
    *out = nullptr;
    int64_t bytes = 0;
    B200_TRY(b200_p2p_bfs_heap_bytes(num_ranks, n_global, &bytes));
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    b200_p2p_bfs *s = new (std::nothrow) b200_p2p_bfs;
    if (!s) return B200_ERR_NOMEM;
    std::memset(s, 0, sizeof(*s));
    s->ctx = ctx;
    s->rank = rank;
    s->P = num_ranks;
    while ((1 << s->log_p) < num_ranks) ++s->log_p;
    s->n_global = n_global;
    s->n_local = n_global / num_ranks;
    s->wl = (uint32_t)(s->n_local / 32);
    const size_t slice = ((size_t)s->wl * 4 + 255) & ~(size_t)255;
    s->off_slice[0] = P2P_CTRL_BYTES;
    s->off_slice[1] = P2P_CTRL_BYTES + slice;
    s->off_known = P2P_CTRL_BYTES + 2 * slice;
    s->off_inbox = s->off_known + (((size_t)(n_global / 8) + 255) & ~(size_t)255);
    s->heap_bytes = (size_t)bytes;
    int st = B200_OK;
    do {
        if ((st = cuda_status(cudaMalloc(&s->heap, s->heap_bytes)))) break;
        if ((st = cuda_status(cudaMemset(s->heap, 0, s->off_inbox)))) break;
        s->known = reinterpret_cast<uint32_t *>(s->heap + s->off_known);   // in the heap: the bitmap exchange reads peers' slices of it
        if ((st = cuda_status(cudaMalloc(&s->full, (size_t)(n_global / 8))))) break;
        if ((st = cuda_status(cudaMalloc(&s->done, sizeof(uint32_t) * (size_t)s->wl)))) break;
        if ((st = cuda_status(cudaMalloc(&s->small, sizeof(SmallShared))))) break;
        if ((st = cuda_status(cudaMemset(s->small, 0, sizeof(SmallShared))))) break;
        if ((st = cuda_status(cudaMalloc(&s->big_rows, sizeof(uint32_t) * (size_t)s->n_local)))) break;
        {
            // the small-level kernel synchronises its grid itself: every CTA must be resident
            int occ = 0;
            if ((st = cuda_status(cudaFuncSetAttribute(p2p_small_levels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM)))) break;
            if ((st = cuda_status(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p2p_small_levels_kernel, SMALL_NT, SMALL_SMEM)))) break;
            if (occ < 1) {
                st = B200_ERR_UNSUPPORTED;
                break;
            }
            s->small_grid = ctx->ws.num_sms;
            const char *eg = getenv("B200_P2P_SMALL_GRID");   // e.g. several ranks sharing one GPU (tests): their grids must fit together
            if (eg && atoi(eg) > 0 && atoi(eg) < s->small_grid) s->small_grid = atoi(eg);
            if ((st = cuda_status(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p2p_pull_levels_kernel, PULL_NT, 0)))) break;
            if (occ < 1) {
                st = B200_ERR_UNSUPPORTED;
                break;
            }
            s->pull_grid = ctx->ws.num_sms * (occ > 4 ? 4 : occ);
            const char *pg = getenv("B200_P2P_PULL_GRID");
            if (pg && atoi(pg) > 0 && atoi(pg) < s->pull_grid) s->pull_grid = atoi(pg);
        }
        if ((st = cuda_status(cudaMalloc(&s->pull_slots, sizeof(PullCounters) * SMALL_SLOTS)))) break;
        if ((st = cuda_status(cudaMemset(s->pull_slots, 0, sizeof(PullCounters) * SMALL_SLOTS)))) break;
        s->absorb_tiles = ((int64_t)s->wl + COMPACT_NT * BITLIST_VT - 1) / (COMPACT_NT * BITLIST_VT) + 1;
        if ((st = cuda_status(cudaMalloc(&s->absorb_status, sizeof(unsigned long long) * (size_t)s->absorb_tiles)))) break;
        if ((st = cuda_status(cudaMemset(s->absorb_status, 0, sizeof(unsigned long long) * (size_t)s->absorb_tiles)))) break;
        if ((st = cuda_status(cudaMalloc(&s->box_counts, sizeof(unsigned long long) * P2P_MAX)))) break;
        if ((st = cuda_status(cudaHostAlloc(&s->h_res, sizeof(unsigned long long) * RES_WORDS, cudaHostAllocMapped)))) break;
        std::memset(s->h_res, 0, sizeof(unsigned long long) * RES_WORDS);
        if ((st = cuda_status(cudaHostGetDevicePointer(&s->d_res, s->h_res, 0)))) break;
        if ((st = cuda_status(cudaEventCreate(&s->ev_run[0])))) break;
        if ((st = cuda_status(cudaEventCreate(&s->ev_run[1])))) break;
        if ((st = ensure_traversal_scratch(ctx, s->n_local))) break;
        if ((st = cuda_status(preload_kernels()))) break;
        if (ipc_handle_out) {
            cudaIpcMemHandle_t h;
            if ((st = cuda_status(cudaIpcGetMemHandle(&h, s->heap)))) break;
            static_assert(sizeof(h) == B200_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
            std::memcpy(ipc_handle_out, &h, sizeof(h));
        }
        if ((st = cuda_status(cudaDeviceSynchronize()))) break;
    } while (0);
    if (st != B200_OK) {
        b200_p2p_bfs_destroy(s);
        return st;
    }
    for (int p = 0; p < P2P_MAX; ++p) s->peers.base[p] = s->heap;   // until connected
    if (heap_base_out) *heap_base_out = s->heap;
    *out = s;
    return B200_OK;
}

int b200_p2p_bfs_connect(b200_p2p_bfs *s, const void *ipc_handles, void *const *peer_bases) {
    if (!s || (s->P > 1 && !ipc_handles && !peer_bases)) return B200_ERR_INVALID;
    B200_CUDA(cudaSetDevice(s->ctx->ws.device));
    for (int p = 0; p < s->P; ++p) {
        if (p == s->rank) {
            s->peers.base[p] = s->heap;
        } else if (peer_bases) {
            if (!peer_bases[p]) return B200_ERR_INVALID;
            // same process: heaps on other devices need peer access enabled once
            cudaPointerAttributes attr;
            B200_CUDA(cudaPointerGetAttributes(&attr, peer_bases[p]));
            if (attr.device != s->ctx->ws.device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_status(e);
                (void)cudaGetLastError();
            }
            s->peers.base[p] = static_cast<unsigned char *>(peer_bases[p]);
        } else {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, static_cast<const unsigned char *>(ipc_handles) + (size_t)p * B200_IPC_HANDLE_BYTES, sizeof(h));
            void *ptr = nullptr;
            B200_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            s->opened[p] = ptr;
            s->peers.base[p] = static_cast<unsigned char *>(ptr);
        }
    }
    s->connected = true;
    return B200_OK;
}

int b200_p2p_bfs_destroy(b200_p2p_bfs *s) {
    if (!s) return B200_OK;
    cudaSetDevice(s->ctx->ws.device);
    cudaStreamSynchronize((cudaStream_t)s->ctx->ws.stream);
    for (int p = 0; p < P2P_MAX; ++p)
        if (s->opened[p]) cudaIpcCloseMemHandle(s->opened[p]);
    if (s->heap) cudaFree(s->heap);
    if (s->full) cudaFree(s->full);
    if (s->done) cudaFree(s->done);
    if (s->small) cudaFree(s->small);
    if (s->big_rows) cudaFree(s->big_rows);
    if (s->pull_slots) cudaFree(s->pull_slots);
    if (s->absorb_status) cudaFree(s->absorb_status);
    if (s->box_counts) cudaFree(s->box_counts);
    if (s->h_res) cudaFreeHost(s->h_res);
    if (s->ev_run[0]) cudaEventDestroy(s->ev_run[0]);
    if (s->ev_run[1]) cudaEventDestroy(s->ev_run[1]);
    if (s->have_level_events)
        for (int i = 0; i <= P2P_TIMED_LEVELS; ++i) cudaEventDestroy(s->ev_level[i]);
    p2p_drop_graph(s);
    if (s->cap_stream) cudaStreamDestroy(s->cap_stream);
    if (s->d_lstate) cudaFree(s->d_lstate);
    if (s->d_trace) cudaFree(s->d_trace);
    delete[] s->h_trace;
    if (s->h_lparams) cudaFreeHost(s->h_lparams);
    if (s->h_lresult) cudaFreeHost(s->h_lresult);
    delete s;
    return B200_OK;
}

int b200_p2p_bfs_set_trace(b200_p2p_bfs *s, int on) {
    if (!s) return B200_ERR_INVALID;
    s->trace_on = on != 0;
    return B200_OK;
}

int b200_p2p_bfs_last_trace(b200_p2p_bfs *s, uint64_t *entries, int64_t capacity, int64_t *count) {
    if (!s || !count || capacity < 0 || (capacity && !entries)) return B200_ERR_INVALID;
    *count = 0;
    if (!s->h_trace) return B200_OK;
    int64_t n = (int64_t)s->h_trace[0];
    if (n > capacity) n = capacity;
    for (int64_t i = 0; i < n; ++i) entries[i] = (uint64_t)s->h_trace[i + 1];
    *count = n;
    return B200_OK;
}

int b200_p2p_bfs_prepare(b200_p2p_bfs *s, const b200_graph *g, int mode, int32_t *d_labels) {
    if (!s || !g || !d_labels || g->n != s->n_local) return B200_ERR_INVALID;
    if (mode != B200_BFS_PUSH && mode != B200_BFS_BEAMER) return B200_ERR_INVALID;
    if (s->P > 1 && !s->connected) return B200_ERR_INVALID;
    b200_ctx *ctx = s->ctx;
    B200_CUDA(cudaSetDevice(ctx->ws.device));
    const bool quad = ctx->adv_impl != B200_ADVANCE_LBS && quad_aligned(g->col_indices);
    if (!quad || ctx->loop_impl != B200_LOOP_GRAPH) return B200_OK;   // host-driven loop: nothing to build
    const int st = p2p_ensure_graph(s, g, d_labels, mode);
    return st == B200_ERR_UNSUPPORTED ? B200_OK : st;                 // run() then takes the host-driven loop
}

int b200_p2p_bfs_run(b200_p2p_bfs *s, const b200_graph *g, int64_t m_global, int32_t src, int mode, float alpha, float beta,
                     int32_t *d_labels, b200_stats *stats, int64_t *sent_per_level) {
    if (!s || !g || !d_labels || src < 0 || src >= s->n_global || g->n != s->n_local || m_global < g->m) return B200_ERR_INVALID;
    if (mode != B200_BFS_PUSH && mode != B200_BFS_BEAMER) return B200_ERR_INVALID;
    if (s->P > 1 && !s->connected) return B200_ERR_INVALID;
    b200_ctx *ctx = s->ctx;
    b200_workspace *ws = &ctx->ws;
    B200_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = ws_stream(ws);
    const int me = s->rank, P = s->P;
    const Partition part{s->log_p, (uint32_t)me, (uint32_t)s->n_local};
    const int64_t n = s->n_global;
    const bool timing = stats && stats->collect_timing;
    if (timing && !s->have_level_events) {
        for (int i = 0; i <= P2P_TIMED_LEVELS; ++i) B200_CUDA(cudaEventCreate(&s->ev_level[i]));
        s->have_level_events = true;
    }
    const int64_t launches0 = ws->launches;
    const uint32_t *pull_off = g->col_offsets ? g->col_offsets : g->row_offsets;
    const int32_t *pull_idx = g->row_indices ? g->row_indices : g->col_indices;
    const bool quad = ctx->adv_impl != B200_ADVANCE_LBS && quad_aligned(g->col_indices);
    Ctrl *my_ctrl = reinterpret_cast<Ctrl *>(s->heap);
    int *my_inbox = reinterpret_cast<int *>(s->heap + s->off_inbox);
    if (alpha <= 0.f) alpha = 15.f;
    if (beta <= 0.f) beta = 18.f;
    if (quad && !timing && ctx->loop_impl == B200_LOOP_GRAPH) {
        // every rank takes the same branch: the choice depends only on settings the ranks share
        const int gs = p2p_run_graph(s, g, m_global, src, mode, alpha, beta, d_labels, stats, sent_per_level);
        if (gs != B200_ERR_UNSUPPORTED) return gs;
    }
    if (stats) stats->level_loop = B200_LOOP_HOST;

    B200_CUDA(cudaEventRecord(s->ev_run[0], st));
    B200_CUDA(cudaMemsetAsync(d_labels, 0xFF, sizeof(int32_t) * (size_t)s->n_local, st));
    B200_CUDA(cudaMemsetAsync(s->known, 0, (size_t)(n / 8), st));
    if (mode == B200_BFS_BEAMER) {   // own rows without in-arcs start "known": pull levels skip them (engine.cuh)
        uint32_t *own = s->known + (size_t)me * s->wl;
        if (g->no_in_arc_bitmap)
            B200_CUDA(cudaMemcpyAsync(own, g->no_in_arc_bitmap, sizeof(uint32_t) * (size_t)s->wl, cudaMemcpyDeviceToDevice, st));
        else
            B200_CUDA(launch_no_in_arc_bitmap(ws, pull_off, s->n_local, own));
    }
    p2p_init_kernel<<<1, 1, 0, st>>>(d_labels, s->known, ctx->frontier[0], src, part);
    ws->launches++;
    B200_CUDA(cudaGetLastError());

    int sel = 0, sb = 0, level = 0;
    bool pull = false;
    int64_t flen_local = part.owner((uint32_t)src) == part.me ? 1 : 0, flen = 1, reached = 1, total_arcs = 0;
    int64_t m_unexplored = m_global;
    int status = B200_OK;

    for (;;) {
        const bool tl = timing && level < P2P_TIMED_LEVELS;
        if (tl && level == 0) B200_CUDA(cudaEventRecord(s->ev_level[0], st));
        B200_CUDA(reset_counters(ws));
        uint32_t *slice_w = reinterpret_cast<uint32_t *>(s->heap + s->off_slice[sb ^ 1]);
        if (!pull) {
            B200_CUDA(cudaMemsetAsync(s->box_counts, 0, sizeof(unsigned long long) * P2P_MAX, st));
            int32_t *next = ctx->frontier[sel ^ 1];
            if (flen_local) {
                RoutedOut r;
                std::memset(&r, 0, sizeof r);
                r.num_dest = P;
                r.count = s->box_counts;
                for (int p = 0; p < P; ++p) {
                    // box[p]: segment `me` of rank p's inbox -- a peer pointer: the flush stores over NVLink
                    r.box[p] = p == me ? next : reinterpret_cast<int *>(s->peers.base[p] + s->off_inbox) + (size_t)me * (size_t)s->n_local;
                    r.capacity[p] = (unsigned long long)s->n_local;
                }
                if (quad) {
                    B200_CUDA(launch_quad_scan(ws, ctx->frontier[sel], (uint32_t)flen_local, g->row_offsets, part.log_p));
                    const QuadArgs a = make_quad_args(ws, ctx->frontier[sel], (uint32_t)flen_local, g->row_offsets, g->col_indices, nullptr, part.log_p);
                    BfsPushPartQ op{s->known, d_labels, level + 1, part};
                    B200_CUDA((launch_quad_advance<OUT_ROUTED, true>(ws, a, op, nullptr, 0ull, &r)));
                } else {
                    B200_CUDA(launch_frontier_scan(ws, ctx->frontier[sel], (uint32_t)flen_local, g->row_offsets, part.log_p));
                    const LbsArgs a = make_lbs_args(ws, ctx->frontier[sel], (uint32_t)flen_local, g->row_offsets, g->col_indices, part.log_p);
                    BfsPushPartOp op{s->known, d_labels, level + 1, part};
                    B200_CUDA((launch_lbs_advance<OUT_ROUTED, true>(ws, a, op, nullptr, 0ull, &r)));
                }
            }
            if (P > 1) {
                p2p_publish_counts_kernel<<<1, 32, 0, st>>>(s->peers, me, P, ++s->epoch, s->box_counts, s->d_res);
                p2p_absorb_kernel<<<ws->num_sms * 2, 256, 0, st>>>(my_inbox, (size_t)s->n_local, my_ctrl->counts, me, P, s->known,
                                                                   d_labels, level + 1, part, next, (unsigned long long)s->n_local,
                                                                   s->box_counts + me, ws->d_counters, g->row_offsets);
                ws->launches += 2;
                B200_CUDA(cudaGetLastError());
            }
            if (mode == B200_BFS_BEAMER) {
                B200_CUDA(cudaMemsetAsync(slice_w, 0, sizeof(uint32_t) * (size_t)s->wl, st));
                p2p_list_to_slice_kernel<<<ws->num_sms, 256, 0, st>>>(next, s->box_counts + me, slice_w, part);
                ws->launches++;
                B200_CUDA(cudaGetLastError());
            }
        } else {
            p2p_gather_or_kernel<<<ws->num_sms * 4, 256, 0, st>>>(s->peers, s->off_slice[sb], s->wl / 4, P,
                                                                  reinterpret_cast<uint4 *>(s->full), reinterpret_cast<uint4 *>(s->known));
            bfs_pull_kernel<256><<<ws->num_sms * 8, 256, 0, st>>>((uint32_t)s->n_local, pull_off, pull_idx, s->full, slice_w,
                                                                  s->known + (size_t)me * s->wl, d_labels, level + 1, ws->d_counters, part,
                                                                  g->first_in_neighbor);
            ws->launches += 2;
            B200_CUDA(cudaGetLastError());
        }
        p2p_stats_kernel<<<1, 32, 0, st>>>(s->peers, me, P, ++s->epoch, (int)(s->stats_seq++ & 1u), ws->d_counters, s->box_counts,
                                           pull ? 0 : 1, (pull || quad) ? B200_CNT_ARCS : B200_CNT_TOTAL, s->d_res);
        ws->launches++;
        B200_CUDA(cudaGetLastError());
        if (tl) B200_CUDA(cudaEventRecord(s->ev_level[level + 1], st));
        B200_CUDA(cudaStreamSynchronize(st));
        if (s->h_res[RES_TIMEOUT]) {
            s->h_res[RES_TIMEOUT] = 0;
            status = B200_ERR_TIMEOUT;
            break;
        }
        if (s->h_res[RES_SUM + ROW_OVERFLOW]) {
            status = B200_ERR_OVERFLOW;
            break;
        }
        const int64_t found = (int64_t)s->h_res[RES_SUM + ROW_NEXT], arcs = (int64_t)s->h_res[RES_SUM + ROW_ARCS];
        const int64_t next_deg = (int64_t)s->h_res[RES_SUM + ROW_DEG], sent = (int64_t)s->h_res[RES_SUM + ROW_SENT];
        const int64_t next_local = (int64_t)s->h_res[RES_OWN + ROW_NEXT];
        if (stats && level < B200_MAX_LEVELS) {
            b200_level_stat *ls = &stats->level[level];
            ls->direction = pull ? 1 : 0;
            ls->frontier_len = flen;
            ls->arcs = arcs;
            ls->discovered = found;
        }
        if (sent_per_level && level < B200_MAX_LEVELS) sent_per_level[level] = sent;
        total_arcs += arcs;
        ++level;
        if (found == 0) break;
        reached += found;
        if (!pull) {
            m_unexplored -= arcs;
            sel ^= 1;
            flen_local = next_local;
            if (mode == B200_BFS_BEAMER && (double)next_deg > (double)m_unexplored / alpha && found > flen) {
                pull = true;
                sb ^= 1;   // the slice written this level is the frontier of the first pull level
            }
        } else {
            sb ^= 1;
            if ((double)found < (double)n / beta && found < flen) {
                // hand the (small) frontier back to push: peers' last pull discoveries become "known",
                // this rank's slice becomes its frontier list
                p2p_gather_or_kernel<<<ws->num_sms * 4, 256, 0, st>>>(s->peers, s->off_slice[sb], s->wl / 4, P,
                                                                      reinterpret_cast<uint4 *>(s->full), reinterpret_cast<uint4 *>(s->known));
                ws->launches++;
                B200_CUDA(cudaGetLastError());
                B200_CUDA(reset_counters(ws));
                B200_CUDA(launch_bitmap_list(ws, BitmapWords{reinterpret_cast<const uint32_t *>(s->heap + s->off_slice[sb])},
                                             PartItem{part}, (uint32_t)s->n_local, ctx->frontier[sel], (unsigned long long)s->n_local,
                                             ws->d_counters + B200_CNT_OUT, ws->d_counters + B200_CNT_OVERFLOW));
                flen_local = next_local;
                pull = false;
            }
        }
        flen = found;
    }
    B200_CUDA(cudaEventRecord(s->ev_run[1], st));
    B200_CUDA(cudaEventSynchronize(s->ev_run[1]));
    if (stats) {
        stats->num_levels = level;
        stats->reached = reached;
        stats->total_arcs = total_arcs;
        stats->launches = ws->launches - launches0;
        B200_CUDA(cudaEventElapsedTime(&stats->device_ms, s->ev_run[0], s->ev_run[1]));
        if (timing) {
            const int L = level < P2P_TIMED_LEVELS ? level : P2P_TIMED_LEVELS;
            for (int l = 0; l < L; ++l) {
                B200_CUDA(cudaEventElapsedTime(&stats->level[l].level_ms, s->ev_level[l], s->ev_level[l + 1]));
                stats->level[l].advance_ms = 0.f;
            }
        }
    }
    return status;
}

}  // extern "C"
