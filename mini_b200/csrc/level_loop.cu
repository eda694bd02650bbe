// level_loop.cu -- the whole BFS as ONE CUDA graph whose level loop is steered from the device.
//
// bfs_enactor_t::enact_pushpull (bfs_enactor.hxx:41-117) returns to the host after every operator:
// two blocking read-backs and ~10 launches per level.  The host-driven loop of engine.cu already cut
// that to one read-back per level, but on a B200 a BFS level of a scale-22 graph is 10-200 us of
// kernel time, so the ~35 us round trip (D2H copy + stream sync + relaunch) was a third of the
// traversal.  Here the traversal is a graph
//
//     memset labels / visited / bitmap 0 -> init -> WHILE(run) { quad scan; quad advance; pull; decide;
//                                                                list -> bitmap; bitmap -> list }
//
// built once per (graph, labels buffer, mode) and replayed by one cudaGraphLaunch: the `decide`
// kernel reads the level's counters on the device, applies exactly the host loop's push / pull / stop
// rules, flips the ping-pong buffers in the device-resident LoopDyn (loop_dyn.cuh), re-arms the
// counters, sets the WHILE handle and the LOOP_RUN_* bits that tell each kernel of the flat body
// whether it has work this level.  (First version: IF/ELSE around push | pull and a SWITCH around the
// transitions -- measured 7 and 4 us per conditional node per level on a B200 against 1.1 us for a
// kernel that returns at once, profiles/microbench/graph_cond.cu, so the body is flat now.)
// Per-level statistics and the result land in mapped pinned memory; the host synchronises once.
// Invariant: frontier bitmap 0 is all-zero during push levels (prologue memset; the bitmap -> list
// compaction clears it again), so the list -> bitmap transition is a plain scatter.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include "b200/operators.cuh"
#include "engine.cuh"
#include "graph_util.cuh"
#include "near_far.cuh"

using namespace b200;

namespace b200 {

struct LoopParams {      // mapped pinned: run parameters, read by the init kernel (not baked into the graph)
    int32_t src, mode;
    float alpha, beta;
    long long m;
    uint32_t epoch0;
    float delta0;        // SSSP: first bucket width of the near-far loop (+inf: the reference's Bellman-Ford iterations)
    unsigned long long *trace;   // B200_LOOP_TRACE: kernel-entry timeline buffer (NULL = off), see loop_dyn.cuh
    uint32_t trace_cap, pad;
};

struct LoopLevelRec {
    int32_t direction, pad;
    long long frontier_len, arcs, discovered;
};

struct LoopResult {      // mapped pinned: written by the decide kernel
    int32_t status, num_levels;
    long long reached, total_arcs, launches;
    LoopLevelRec level[B200_MAX_LEVELS];
};

struct LoopState {       // device
    LoopDyn dyn;
    int32_t level, pull, mode, pad;
    float alpha, beta;
    long long n, m_unexplored, unvisited, reached, total_arcs, flen, launches;
};

struct LevelLoop {
    LoopState *d_state;
    LoopParams *h_params, *d_params;
    LoopResult *h_result, *d_result;
    unsigned long long *d_trace;   // [LOOP_TRACE_CAP], allocated by the first traced run
    cudaStream_t cap_stream;
    // one cached graph per primitive (rebuilt when its key changes)
    struct Slot {
        cudaGraph_t graph;
        cudaGraphExec_t exec;
        const void *k_offsets, *k_indices, *k_labels, *k_scratch, *k_iso, *k_first, *k_weights, *k_pull_offsets, *k_pull_indices;
        int64_t k_n;
        uint64_t k_gen;
        int k_mode;
        bool k_wc;
    } slot[2];
    int failed;          // status of the last failed build (the caller falls back to the host loop)
};
enum { SLOT_BFS = 0, SLOT_SSSP = 1, MODE_SSSP = 100, LOOP_TRACE_CAP = 4096 };

}  // namespace b200

namespace {

#ifndef B200_LOOP_UNROLL
#define B200_LOOP_UNROLL 3
#endif
inline int loop_unroll() {   // levels per pass of the WHILE body (the B200_LOOP_UNROLL environment variable overrides the build's default)
    const char *e = getenv("B200_LOOP_UNROLL");
    const int v = e ? atoi(e) : B200_LOOP_UNROLL;
    return v < 1 ? 1 : v > 8 ? 8 : v;
}

constexpr int DYN_SCAN_CTAS_PER_SM = 8;   // 256-thread CTAs: the whole SM, as many tiles in flight as a count-sized grid has

struct FrontierQuadsDyn {   // FrontierQuads over the device-selected frontier list
    const LoopDyn *dyn;
    const uint32_t *offsets;
    uint2 *rows;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        const int v = dyn->in[i];
        uint32_t b = 0, e = 0;
        if (v >= 0) {
            b = __ldg(offsets + v);
            e = __ldg(offsets + v + 1);
        }
        (dyn->rows_in ? dyn->rows_in : rows)[i] = make_uint2(b, e);
        return e > b ? ((e - 1) >> 2) - (b >> 2) + 1u : 0u;
    }
};

struct ScanPairs {   // the two (scan, row bounds) buffer pairs of the workspace; all NULL = re-scan every level
    uint32_t *sc0, *sc1;
    uint2 *rw0, *rw1;
};

__global__ void sssp_loop_fill_kernel(float *dist, int32_t *stamp, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        dist[i] = 3.402823466e+38f;
        stamp[i] = -1;
    }
}

// bfs_problem_t ctor (bfs_problem.hxx:38-42) + init_frontier (bfs_enactor.hxx:34-38) + the loop state;
// SSSP: labels holds float distances (sssp_problem.hxx:44-46: source 0.0f)
template <bool SSSP>
__global__ void loop_init_kernel(const LoopParams *p, LoopState *s, int32_t *labels, uint32_t *visited, int32_t *f0,
                                 int32_t *f1, long long n, unsigned long long *counters, unsigned int *tile_counters,
                                 ScanPairs sp, NearFar *nf) {
    const int src = p->src;
    labels[src] = 0;                         // depth 0 / +0.0f
    if (!SSSP) visited[src >> 5] |= 1u << (src & 31);
    f0[0] = src;
    s->dyn.in = f0;
    s->dyn.out = f1;
    s->dyn.len = 1u;
    s->dyn.epoch = p->epoch0 & 0x3FFFFFFFu;
    s->dyn.next_label = 1;
    s->dyn.bsel = 0u;
    s->dyn.run = LOOP_RUN_PUSH | LOOP_RUN_SCAN;
    s->dyn.scanned_in = sp.sc0;
    s->dyn.rows_in = sp.rw0;
    s->dyn.scanned_out = sp.sc1;
    s->dyn.rows_out = sp.rw1;
    s->dyn.trace = p->trace;
    s->dyn.trace_cap = p->trace_cap;
    if (p->trace) p->trace[0] = 0ull;
    s->level = 0;
    s->pull = 0;
    s->mode = p->mode;
    s->alpha = p->alpha;
    s->beta = p->beta;
    s->n = n;
    s->m_unexplored = p->m;
    s->unvisited = n - 1;
    s->reached = 1;
    s->total_arcs = 0;
    s->flen = 1;
    s->launches = 1;
    for (int i = 0; i < B200_NUM_COUNTERS; ++i) counters[i] = 0ull;
    tile_counters[0] = 0u;
    tile_counters[1] = 0u;
    if (SSSP) near_far_reset(nf, p->delta0, p->m);
}

// The host loop's per-level bookkeeping (engine.cu b200_bfs_run), on the device.
__global__ void loop_decide_kernel(LoopState *s, unsigned long long *counters, unsigned int *tile_counters, LoopResult *res,
                                   cudaGraphConditionalHandle h_while, NearFar *nf) {
    if (!(s->dyn.run & (LOOP_RUN_PUSH | LOOP_RUN_PULL))) return;   // no level ran in this pass of the (unrolled) body: the
                                                                   // traversal is over, or the near-far passes just ran
    loop_trace(&s->dyn, 9);
    const bool was_pull = s->pull != 0;
    const bool work_create = s->dyn.scanned_in != nullptr;
    long long found = (long long)counters[B200_CNT_OUT];
    unsigned long long next_quads = 0ull;
    if (work_create && !was_pull) {   // the advance packed (vertices emitted << 32) | quads created
        next_quads = (unsigned long long)found & 0xffffffffull;
        found = (long long)((unsigned long long)found >> 32);
    }
    const long long arcs = (long long)counters[B200_CNT_ARCS];
    const long long next_deg = (long long)counters[B200_CNT_AUX];
    const bool overflow = counters[B200_CNT_OVERFLOW] != 0ull;
    bool need_scan = !work_create;
    bool pull = was_pull;
    int level = s->level;
    if (level < B200_MAX_LEVELS) {
        LoopLevelRec *r = &res->level[level];
        r->direction = was_pull ? 1 : 0;
        r->frontier_len = was_pull ? s->unvisited : s->flen;
        r->arcs = arcs;
        r->discovered = found;
    }
    s->total_arcs += arcs;
    s->launches += s->mode == MODE_SSSP ? 5 : s->mode == B200_BFS_PUSH ? 3 : 6;   // kernel nodes of the flat body (some return at once)
    if (nf) nf->bucket_arcs += arcs;
    ++level;
    s->level = level;
    bool done = false, take = false;
    uint32_t trans = 0u;
    int status = B200_OK;
    if (overflow) {
        status = B200_ERR_OVERFLOW;
        done = true;
    } else if (found == 0) {
        if (nf) take = true;    // the near frontier ran dry: sssp_pending_min_kernel ends the traversal or opens a bucket
        else done = true;
    } else {
        const long long flen = s->flen, n = s->n;
        s->reached += found;
        s->unvisited -= found;
        if (!was_pull) {
            s->m_unexplored -= arcs;
            const int *t = s->dyn.in;
            s->dyn.in = s->dyn.out;
            s->dyn.out = const_cast<int *>(t);
            if (work_create) {   // the scan and row bounds the advance created become this level's
                uint32_t *ts = s->dyn.scanned_in;
                s->dyn.scanned_in = s->dyn.scanned_out;
                s->dyn.scanned_out = ts;
                uint2 *tr = s->dyn.rows_in;
                s->dyn.rows_in = s->dyn.rows_out;
                s->dyn.rows_out = tr;
            }
            if (s->mode == B200_BFS_REF_ALPHA) {
                if ((float)s->unvisited < (float)found * s->alpha) pull = true;   // bfs_enactor.hxx:68
            } else if (s->mode == B200_BFS_BEAMER) {
                if ((double)next_deg > (double)s->m_unexplored / s->alpha && found > flen) pull = true;
            }
        } else {
            s->dyn.bsel ^= 1u;
            if (s->mode == B200_BFS_BEAMER && (double)found < (double)n / s->beta && found < flen) pull = false;
        }
        s->flen = found;
        s->dyn.len = (uint32_t)found;
        if (pull && !was_pull) {
            trans = LOOP_RUN_TO_PULL;   // sparse_to_dense_kernel (advance.hxx:69-84) into bitmap 0
            s->dyn.bsel = 0u;
        } else if (!pull && was_pull) {
            trans = LOOP_RUN_TO_PUSH;   // bitmap -> list for the push levels that finish the traversal
            need_scan = true;           // ... a list nobody has scanned
        }
    }
    s->pull = pull ? 1 : 0;
    s->dyn.next_label = level + 1;
    s->dyn.epoch = (s->dyn.epoch + 2u) & 0x3FFFFFFFu;
    for (int i = 0; i < B200_NUM_COUNTERS; ++i) counters[i] = 0ull;
    if (!need_scan) counters[B200_CNT_TOTAL] = next_quads;   // what the scan kernel would have left there
    tile_counters[0] = 0u;
    tile_counters[1] = 0u;
    if (done) {
        res->status = status;
        res->num_levels = level;
        res->reached = s->reached;
        res->total_arcs = s->total_arcs;
        res->launches = s->launches;
        __threadfence_system();
    }
    loop_trace(&s->dyn, 64u + (unsigned)((level - 1) & 63));   // level closed
    s->dyn.run = done ? 0u : take ? (uint32_t)LOOP_RUN_TAKE : ((pull ? LOOP_RUN_PULL : LOOP_RUN_PUSH | (need_scan ? LOOP_RUN_SCAN : 0u)) | trans);
    cudaGraphSetConditional(h_while, done ? 0u : 1u);
}

// near_far.cuh's passes inside the graph: they run when the decide step found the near frontier empty; the min pass
// ends the traversal (what the decide step does for BFS) or opens a bucket, the take pass hands the level loop its
// next frontier -- a list nobody has scanned.
struct NearFarGraphHook {
    LoopState *s;
    LoopResult *res;
    cudaGraphConditionalHandle h_while;
    __device__ __forceinline__ bool enabled() const { return (s->dyn.run & LOOP_RUN_TAKE) != 0u; }
    __device__ __forceinline__ void trace(unsigned id) const { loop_trace(&s->dyn, id); }
    __device__ __forceinline__ void pre(NearFar *) const {}
    __device__ __forceinline__ void on_done(NearFar *) const {
        res->status = B200_OK;
        res->num_levels = s->level;
        res->reached = s->reached;
        res->total_arcs = s->total_arcs;
        res->launches = s->launches;
        __threadfence_system();
        s->dyn.run = 0u;
        cudaGraphSetConditional(h_while, 0u);
    }
    __device__ __forceinline__ int *out() const { return const_cast<int *>(s->dyn.in); }
    __device__ __forceinline__ void on_taken(NearFar *, uint32_t count) const {
        s->flen = count;
        s->dyn.len = count;
        s->dyn.run = LOOP_RUN_PUSH | LOOP_RUN_SCAN;
    }
};

void drop_slot(LevelLoop::Slot *S) {
    if (S->exec) cudaGraphExecDestroy(S->exec);
    if (S->graph) cudaGraphDestroy(S->graph);
    std::memset(S, 0, sizeof *S);
}
void drop_graph(LevelLoop *L) {
    drop_slot(&L->slot[SLOT_BFS]);
    drop_slot(&L->slot[SLOT_SSSP]);
}

// mode: a B200_BFS_* mode (d_labels = int32 labels) or MODE_SSSP (d_labels = float distances, reinterpreted)
int build_graph(b200_ctx *ctx, const b200_graph *g, int32_t *d_labels, int mode, const uint32_t *iso, bool work_create) {
    LevelLoop *L = ctx->loop;
    LevelLoop::Slot *S = &L->slot[mode == MODE_SSSP ? SLOT_SSSP : SLOT_BFS];
    const bool sssp = mode == MODE_SSSP;
    const bool has_pull = !sssp && mode != B200_BFS_PUSH;
    b200_workspace *ws = &ctx->ws;
    const int64_t n = g->n;
    const size_t words = (size_t)((n + 31) / 32);
    const bool deg = mode == B200_BFS_BEAMER;
    const uint32_t *pull_off = g->col_offsets ? g->col_offsets : g->row_offsets;
    const int32_t *pull_idx = g->row_indices ? g->row_indices : g->col_indices;
    cudaStream_t cs = L->cap_stream;
    void *const user_stream = ws->stream;
    const int64_t launches0 = ws->launches;
    int st = B200_OK;
    bool capturing = false;
    cudaGraph_t G = nullptr, body = nullptr;
    cudaGraphConditionalHandle h_while;
    cudaGraphNode_t deps[8], n_while;
    size_t ndeps = 0;
    cudaGraphNodeParams np_w = {};
    const LoopDyn *dyn = &L->d_state->dyn;
    ScanPairs sp = {nullptr, nullptr, nullptr, nullptr};
    if (work_create)
        sp = ScanPairs{ws->d_scanned, ws->d_scanned2, reinterpret_cast<uint2 *>(ws->d_rows), reinterpret_cast<uint2 *>(ws->d_rows2)};

    drop_slot(S);
    ws->stream = (void *)cs;   // the launchers below record into the capture stream
    LL_CUDA(cudaGraphCreate(&G, 0));
    LL_CUDA(cudaGraphConditionalHandleCreate(&h_while, G, 1, cudaGraphCondAssignDefault));

    // ---- prologue
    LL_CUDA(cudaStreamBeginCaptureToGraph(cs, G, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    capturing = true;
    if (sssp) {
        // sssp_problem_t ctor (sssp_problem.hxx:44-49) + init_frontier (sssp_enactor.hxx:33-37)
        sssp_loop_fill_kernel<<<ws->num_sms * 4, 256, 0, cs>>>(reinterpret_cast<float *>(d_labels), ctx->stamp, (unsigned long long)n);
        LL_CUDA(cudaGetLastError());
        loop_init_kernel<true><<<1, 1, 0, cs>>>(L->d_params, L->d_state, d_labels, nullptr, ctx->frontier[0], ctx->frontier[1],
                                                (long long)n, ws->d_counters, ws->d_tile_counter, sp, ctx->d_nf);
    } else {
        LL_CUDA(cudaMemsetAsync(d_labels, 0xFF, sizeof(int32_t) * (size_t)n, cs));
        LL_CUDA(cudaMemsetAsync(ctx->bm_visited, 0, sizeof(uint32_t) * words, cs));
        if (mode != B200_BFS_PUSH) LL_CUDA(cudaMemsetAsync(ctx->bm_frontier[0], 0, sizeof(uint32_t) * words, cs));   // frontier bitmap 0 starts clean
        loop_init_kernel<false><<<1, 1, 0, cs>>>(L->d_params, L->d_state, d_labels, ctx->bm_visited, ctx->frontier[0],
                                                 ctx->frontier[1], (long long)n, ws->d_counters, ws->d_tile_counter, sp, nullptr);
    }
    LL_CUDA(cudaGetLastError());
    capturing = false;
    if ((st = end_capture(cs, deps, &ndeps, 8)) != B200_OK) goto fail;

    // ---- WHILE(run)
    np_w.type = cudaGraphNodeTypeConditional;
    np_w.conditional.handle = h_while;
    np_w.conditional.type = cudaGraphCondTypeWhile;
    np_w.conditional.size = 1;
    LL_CUDA(cudaGraphAddNode(&n_while, G, deps, ndeps, &np_w));
    body = np_w.conditional.phGraph_out[0];

    // ---- the flat body: every kernel returns at once unless its LOOP_RUN_* bit is set
    LL_CUDA(cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    capturing = true;
    // The body holds loop_unroll() levels: an iteration of the WHILE node costs ~4 us between the last kernel of one
    // pass and the first of the next (B200_LOOP_TRACE), an idle kernel 1.1-1.8 us -- once the traversal is done the
    // rest of the pass is idle kernels.
    for (int rep = 0; rep < loop_unroll(); ++rep) {
    {
        // push level: quad scan + quad advance (fused uniquify filter)
        const int64_t max_tiles = (n + SCAN_NT * SCAN_VT - 1) / (SCAN_NT * SCAN_VT);
        int64_t grid = (int64_t)ws->num_sms * DYN_SCAN_CTAS_PER_SM;
        if (grid > max_tiles) grid = max_tiles;
        FrontierQuadsDyn fn{dyn, g->row_offsets, reinterpret_cast<uint2 *>(ws->d_rows)};
        scan_sizes_dyn_kernel<SCAN_NT, SCAN_VT><<<(unsigned)grid, SCAN_NT, 0, cs>>>(
            fn, dyn, (uint32_t)LOOP_RUN_SCAN, ws->d_scanned, ws->d_status, ws->d_tile_counter, ws->d_counters + B200_CNT_TOTAL);
        LL_CUDA(cudaGetLastError());
        QuadArgs a = make_quad_args(ws, ctx->frontier[0], 0u, g->row_offsets, g->col_indices, sssp ? g->col_values : nullptr);
        a.dyn = dyn;
        if (sssp) {
            SsspRelaxQDyn op{{reinterpret_cast<float *>(d_labels), nullptr, ctx->stamp, 0, ctx->d_nf}, dyn};   // preds: one exact pass at the end
            LL_CUDA((launch_quad_advance<OUT_COMPACT, false>(ws, a, op, ctx->frontier[1], (unsigned long long)n)));
        } else {
            BfsPushQDyn op{{ctx->bm_visited, d_labels, 0}, dyn};
            if (deg) LL_CUDA((launch_quad_advance<OUT_COMPACT, true>(ws, a, op, ctx->frontier[1], (unsigned long long)n)));
            else LL_CUDA((launch_quad_advance<OUT_COMPACT, false>(ws, a, op, ctx->frontier[1], (unsigned long long)n)));
        }
    }
    if (has_pull) {
        // pull level
        bfs_pull_dyn_kernel<256><<<ws->num_sms * 8, 256, 0, cs>>>((uint32_t)n, pull_off, pull_idx, ctx->bm_frontier[0],
                                                                  ctx->bm_frontier[1], ctx->bm_visited, d_labels, dyn,
                                                                  ws->d_counters, Partition{0, 0, (uint32_t)n},
                                                                  g->first_in_neighbor);
        LL_CUDA(cudaGetLastError());
    }
    loop_decide_kernel<<<1, 1, 0, cs>>>(L->d_state, ws->d_counters, ws->d_tile_counter, L->d_result, h_while,
                                        sssp ? ctx->d_nf : nullptr);
    LL_CUDA(cudaGetLastError());
    if (sssp) {
        const NearFarGraphHook hook{L->d_state, L->d_result, h_while};
        const unsigned grid = nf_grid(ws->num_sms, n);
        sssp_pending_min_kernel<<<grid, NF_NT, 0, cs>>>(reinterpret_cast<const float *>(d_labels), (uint32_t)n, ctx->d_nf, hook);
        LL_CUDA(cudaGetLastError());
        sssp_take_kernel<<<grid, NF_NT, 0, cs>>>(reinterpret_cast<const float *>(d_labels), (uint32_t)n, ctx->d_nf, hook);
        LL_CUDA(cudaGetLastError());
    }
    if (has_pull) {
        // transitions: list -> bitmap 0 (all-zero by invariant) | bitmap -> list (+ clears bitmap 0)
        // (+ vertices without in-arcs count as visited from the switch on, engine.cuh)
        sparse_to_bitmap_dyn_kernel<<<ws->num_sms * 4, 256, 0, cs>>>(dyn, ctx->bm_frontier[0], pull_off, (uint32_t)n, iso, ctx->bm_visited);
        LL_CUDA(cudaGetLastError());
        const int64_t num_words = (n + 31) / 32;
        const int64_t max_tiles = (num_words + COMPACT_NT * BITLIST_VT - 1) / (COMPACT_NT * BITLIST_VT);
        int64_t grid = (int64_t)ws->num_sms * DYN_SCAN_CTAS_PER_SM;
        if (grid > max_tiles) grid = max_tiles;
        bitmap_list_dyn_kernel<COMPACT_NT, BITLIST_VT><<<(unsigned)grid, COMPACT_NT, 0, cs>>>(
            BitmapWordsDyn{dyn, ctx->bm_frontier[0], ctx->bm_frontier[1]}, IdentityItem{}, (uint32_t)num_words, dyn,
            (uint32_t)LOOP_RUN_TO_PUSH, (unsigned long long)n, ws->d_status, ws->d_tile_counter + 1,
            ws->d_counters + B200_CNT_AUX2, ws->d_counters + B200_CNT_OVERFLOW, ctx->bm_frontier[0]);
        LL_CUDA(cudaGetLastError());
    }
    }   // loop_unroll()
    capturing = false;
    if ((st = end_capture(cs, deps, &ndeps, 8)) != B200_OK) goto fail;

    LL_CUDA(cudaGraphInstantiate(&S->exec, G, 0));
    S->graph = G;
    S->k_offsets = g->row_offsets;
    S->k_indices = g->col_indices;
    S->k_weights = g->col_values;
    S->k_labels = d_labels;
    S->k_scratch = ctx->frontier[0];
    S->k_iso = iso;
    S->k_first = g->first_in_neighbor;
    S->k_pull_offsets = pull_off;
    S->k_pull_indices = pull_idx;
    S->k_n = n;
    S->k_gen = ctx->scratch_gen;
    S->k_mode = mode;
    S->k_wc = work_create;
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

}  // namespace

namespace b200 {

void level_loop_destroy(b200_ctx *ctx) {
    LevelLoop *L = ctx->loop;
    if (!L) return;
    drop_graph(L);
    if (L->cap_stream) cudaStreamDestroy(L->cap_stream);
    if (L->d_state) cudaFree(L->d_state);
    if (L->d_trace) cudaFree(L->d_trace);
    if (L->h_params) cudaFreeHost(L->h_params);
    if (L->h_result) cudaFreeHost(L->h_result);
    delete L;
    ctx->loop = nullptr;
}

// The traversal scratch was reallocated: a cached graph holds dangling pointers.
void level_loop_invalidate(b200_ctx *ctx) {
    if (ctx->loop) drop_graph(ctx->loop);
}

static int level_loop_get(b200_ctx *ctx) {
    if (ctx->loop) return B200_OK;
    LevelLoop *L = new (std::nothrow) LevelLoop;
    if (!L) return B200_ERR_NOMEM;
    std::memset(L, 0, sizeof *L);
    ctx->loop = L;
    B200_CUDA(cudaStreamCreateWithFlags(&L->cap_stream, cudaStreamNonBlocking));
    B200_CUDA(cudaMalloc(&L->d_state, sizeof(LoopState)));
    B200_CUDA(cudaHostAlloc(&L->h_params, sizeof(LoopParams), cudaHostAllocMapped));
    B200_CUDA(cudaHostAlloc(&L->h_result, sizeof(LoopResult), cudaHostAllocMapped));
    B200_CUDA(cudaHostGetDevicePointer(&L->d_params, L->h_params, 0));
    B200_CUDA(cudaHostGetDevicePointer(&L->d_result, L->h_result, 0));
    return B200_OK;
}

// b200_bfs_run / b200_sssp_run through the graph.  Returns B200_ERR_UNSUPPORTED when the graph cannot be built on this
// driver (the caller then runs the host-driven loop and reports level_loop = host in the stats).
static int run_graph(b200_ctx *ctx, const b200_graph *g, int32_t src, int mode, float alpha, float beta, int32_t *d_labels,
                     b200_stats *stats, float delta0 = 0.f) {
    B200_TRY(level_loop_get(ctx));
    LevelLoop *L = ctx->loop;
    LevelLoop::Slot *S = &L->slot[mode == MODE_SSSP ? SLOT_SSSP : SLOT_BFS];
    b200_workspace *ws = &ctx->ws;
    cudaStream_t st = ws_stream(ws);
    if (L->failed) return B200_ERR_UNSUPPORTED;
    const uint32_t *iso = (mode != B200_BFS_PUSH && mode != MODE_SSSP) ? g->no_in_arc_bitmap : nullptr;
    const bool work_create = ctx->adv_impl == B200_ADVANCE_QUAD && g->n <= WORK_CREATE_MAX_N;   // (else: scan before every level)
    if (!S->exec || S->k_offsets != g->row_offsets || S->k_indices != g->col_indices || S->k_labels != d_labels ||
        S->k_scratch != ctx->frontier[0] || S->k_n != g->n || S->k_mode != mode || S->k_iso != iso ||
        S->k_first != g->first_in_neighbor || S->k_weights != g->col_values || S->k_wc != work_create ||
        S->k_pull_offsets != (g->col_offsets ? g->col_offsets : g->row_offsets) ||
        S->k_pull_indices != (g->row_indices ? g->row_indices : g->col_indices) ||
        S->k_gen != ctx->scratch_gen) {
        B200_CUDA(cudaStreamSynchronize(st));   // a replay of the old graph may still be running
        const int bs = build_graph(ctx, g, d_labels, mode, iso, work_create);
        if (bs != B200_OK) {
            L->failed = bs;
            return B200_ERR_UNSUPPORTED;
        }
    }
    // look-back tags: the device consumes two per level; keep the host's counter ahead of them
    if (ws->epoch > 0x3FFFFFFFu - 4u * (B200_MAX_LEVELS + 2)) {
        B200_CUDA(cudaMemsetAsync(ws->d_status, 0, sizeof(unsigned long long) * (size_t)ws->status_tiles, st));
        ws->epoch = 0;
    }
    L->h_params->src = src;
    L->h_params->mode = mode;
    L->h_params->alpha = alpha;
    L->h_params->beta = beta;
    L->h_params->m = g->m;
    L->h_params->epoch0 = ws->epoch + 1;
    L->h_params->delta0 = delta0;
    const bool want_trace = getenv("B200_LOOP_TRACE") != nullptr;   // (read per run)
    if (want_trace && !L->d_trace) B200_CUDA(cudaMalloc(&L->d_trace, sizeof(unsigned long long) * LOOP_TRACE_CAP));
    L->h_params->trace = want_trace ? L->d_trace : nullptr;
    L->h_params->trace_cap = want_trace ? (uint32_t)LOOP_TRACE_CAP : 0u;
    L->h_result->status = -1;
    L->h_result->num_levels = 0;
    B200_CUDA(cudaEventRecord(ctx->ev_run[0], st));
    B200_CUDA(cudaGraphLaunch(S->exec, st));
    B200_CUDA(cudaEventRecord(ctx->ev_run[1], st));
    B200_CUDA(cudaEventSynchronize(ctx->ev_run[1]));
    if (want_trace) {   // same line format as the peer-memory BFS (profiles/format_trace.py)
        static thread_local unsigned long long h_trace[LOOP_TRACE_CAP];
        B200_CUDA(cudaMemcpy(h_trace, L->d_trace, sizeof h_trace, cudaMemcpyDeviceToHost));
        const unsigned long long cnt = h_trace[0] < LOOP_TRACE_CAP - 1 ? h_trace[0] : LOOP_TRACE_CAP - 1;
        fprintf(stderr, "B200_LOOP_TRACE rank0 %llu entries (us since first, kernel id):", cnt);
        for (unsigned long long i = 1; i <= cnt; ++i)
            fprintf(stderr, " %.1f:%llu", (double)((h_trace[i] >> 8) - (h_trace[1] >> 8)) * 1e-3, h_trace[i] & 0xffull);
        fprintf(stderr, "\n");
    }
    const LoopResult *r = L->h_result;
    if (r->status < 0) return B200_ERR_CUDA;    // the graph ended without the decide kernel reporting
    const int levels = r->num_levels;
    const unsigned used = 2u * (unsigned)levels + 2u;
    if (used > 0x3FFFFFFFu - ws->epoch - 2u) {  // a traversal with ~2^29 levels: retire every tag
        B200_CUDA(cudaMemsetAsync(ws->d_status, 0, sizeof(unsigned long long) * (size_t)ws->status_tiles, st));
        ws->epoch = 0;
    } else {
        ws->epoch += used;
    }
    ws->launches += r->launches;
    if (stats) {
        stats->num_levels = levels;
        stats->reached = r->reached;
        stats->total_arcs = r->total_arcs;
        stats->launches = r->launches;
        stats->level_loop = B200_LOOP_GRAPH;
        B200_CUDA(cudaEventElapsedTime(&stats->device_ms, ctx->ev_run[0], ctx->ev_run[1]));
        const int nl = levels < B200_MAX_LEVELS ? levels : B200_MAX_LEVELS;
        for (int l = 0; l < nl; ++l) {
            b200_level_stat *ls = &stats->level[l];
            ls->direction = r->level[l].direction;
            ls->frontier_len = r->level[l].frontier_len;
            ls->arcs = r->level[l].arcs;
            ls->discovered = r->level[l].discovered;
            ls->advance_ms = 0.f;
            ls->level_ms = 0.f;
        }
    }
    return r->status;
}

int bfs_run_graph(b200_ctx *ctx, const b200_graph *g, int32_t src, int mode, float alpha, float beta, int32_t *d_labels,
                  b200_stats *stats) {
    return run_graph(ctx, g, src, mode, alpha, beta, d_labels, stats);
}

// The frontier iterations of b200_sssp_run (sssp_enactor.hxx:40-72; near-far ordering, near_far.cuh) through the graph:
// distances final at return; the caller adds the predecessor pass.  ev_run[0] is recorded at the start of the traversal.
int sssp_run_graph(b200_ctx *ctx, const b200_graph *g, int32_t src, float *d_dist, float delta0, b200_stats *stats) {
    return run_graph(ctx, g, src, MODE_SSSP, 0.f, 0.f, reinterpret_cast<int32_t *>(d_dist), stats, delta0);
}

}  // namespace b200
