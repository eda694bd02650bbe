/*
 * oracle.c -- CPU restatement of the mini-gunrock frontier-traversal hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (mini_b200/, include/) may
 * call, link or import this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the *checker* and as the reported CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 *   (a) against the known-answer vectors in tests/golden/ that were produced by
 *       running the reference's own CPU code (bfs_problem_t::cpu,
 *       sssp_problem_t::cpu, load_graph) in this container
 *       (tests/golden/make_golden.py + oracle/ref_shim.cu), and
 *   (b) live against oracle/_ref/libref_cpu.so (the reference's unmodified
 *       headers compiled where they lie) on seeded random graphs.
 *   Exception: neighborhood_reduce / PR have no validation in the reference
 *   (gunrock/tests/pr/test_pr.cu:36-40 only prints) => "parity unpinned by the
 *   reference's tests" for those two; they restate neighborhood.hxx:47-58 and
 *   pr_functor.hxx:11-29 in fp64/fp32 and are cross-checked on the GPU box
 *   against the operator semantics only.
 *
 * All citations are relative to /root/reference/.
 * Conventions: int64 row offsets (the reference is int32-only and cannot hold
 * RMAT scale-26's 2^31 arcs, graph.hxx:19-26), int32 vertex ids.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Synthetic input: counter-based RMAT (not from the reference, */
/* which has no generator).  The product's GPU generator must be bit-identical */
/* (tests/test_rmat_gpu.py).                                                    */
/* ------------------------------------------------------------------------- */

static inline uint64_t orc_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* (a,b,c,d) = (0.57,0.19,0.19,0.05) as exact 32-bit integer thresholds */
#define ORC_T_A   0x91EB851Eu /* floor(0.57 * 2^32) */
#define ORC_T_AB  0xC28F5C28u /* floor(0.76 * 2^32) */
#define ORC_T_ABC 0xF3333333u /* floor(0.95 * 2^32) */

static inline void orc_rmat_pair(uint64_t key, uint64_t e, int scale, uint32_t *pu, uint32_t *pv) {
    uint32_t u = 0, v = 0;
    for (int lvl = 0; lvl < scale; lvl += 2) {
        uint64_t h = orc_mix64(key + ((e << 6) | (uint64_t)(lvl >> 1)));
        uint32_t r = (uint32_t)(h >> 32);
        for (int k = 0; k < 2 && lvl + k < scale; ++k) {
            uint32_t ub = (r >= ORC_T_AB);
            uint32_t vb = (r >= ORC_T_A && r < ORC_T_AB) || (r >= ORC_T_ABC);
            u = (u << 1) | ub;
            v = (v << 1) | vb;
            r = (uint32_t)h;
        }
    }
    *pu = u; *pv = v;
}

ORC_API uint64_t orc_rmat_key(uint64_t seed) { return orc_mix64(seed ^ 0x243F6A8885A308D3ull); }

/* npairs = edge_factor << scale generated (u,v) pairs */
ORC_API void orc_rmat_pairs(int scale, int64_t npairs, uint64_t seed, int32_t *src, int32_t *dst) {
    uint64_t key = orc_rmat_key(seed);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < npairs; ++e) {
        uint32_t u, v;
        orc_rmat_pair(key, (uint64_t)e, scale, &u, &v);
        src[e] = (int32_t)u; dst[e] = (int32_t)v;
    }
}

/* integer weight in [1,64] per undirected pair {u,v}; identical for both
 * directions and for duplicate pairs (so sorted CSR is unique). */
static inline float orc_pair_weight(uint64_t wkey, uint32_t u, uint32_t v) {
    uint32_t lo = u < v ? u : v, hi = u < v ? v : u;
    uint64_t h = orc_mix64(wkey ^ (((uint64_t)lo << 32) | hi));
    return (float)(1 + (int)((h >> 40) & 63));
}
ORC_API uint64_t orc_weight_key(uint64_t wseed) { return orc_mix64(wseed ^ 0x13198A2E03707344ull); }

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* Build CSR the way graph.hxx:159-172 does from (src,dst)-sorted tuples:
 * duplicates and self loops kept, arcs ordered by (src,dst), trailing empty
 * rows get offset m.  If symmetrize != 0 every pair also contributes its
 * reverse (graph.hxx:130-137).  offsets has n+1 entries, indices m entries with
 * m = npairs * (symmetrize ? 2 : 1).  weights may be NULL. */
ORC_API void orc_build_csr(int64_t n, int64_t npairs, const int32_t *src, const int32_t *dst,
                           int symmetrize, int64_t *offsets, int32_t *indices,
                           float *weights, uint64_t wseed) {
    int64_t *cursor = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    for (int64_t e = 0; e < npairs; ++e) {
        cursor[src[e] + 1]++;
        if (symmetrize) cursor[dst[e] + 1]++;
    }
    offsets[0] = 0;
    for (int64_t v = 0; v < n; ++v) offsets[v + 1] = offsets[v] + cursor[v + 1];
    for (int64_t v = 0; v < n; ++v) cursor[v] = offsets[v];
    for (int64_t e = 0; e < npairs; ++e) {
        indices[cursor[src[e]]++] = dst[e];
        if (symmetrize) indices[cursor[dst[e]]++] = src[e];
    }
    free(cursor);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t v = 0; v < n; ++v) {
        int64_t b = offsets[v], e = offsets[v + 1];
        if (e - b > 1) qsort(indices + b, (size_t)(e - b), sizeof(int32_t), cmp_i32);
    }
    if (weights) {
        uint64_t wkey = orc_weight_key(wseed);
#pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t v = 0; v < n; ++v)
            for (int64_t i = offsets[v]; i < offsets[v + 1]; ++i)
                weights[i] = orc_pair_weight(wkey, (uint32_t)v, (uint32_t)indices[i]);
    }
}

/* ------------------------------------------------------------------------- */
/* BFS depths: gunrock/src/bfs/bfs_problem.hxx:52-72 (serial queue BFS;        */
/* labels -1 = unreached, labels[src] = 0; relabel if unvisited or shorter).    */
/* labels must be pre-filled with -1 by the caller exactly as                   */
/* tests/bfs/test_bfs.cu:44 does.                                               */
/* ------------------------------------------------------------------------- */
ORC_API void orc_bfs(int64_t n, const int64_t *row_offsets, const int32_t *col_indices,
                     int32_t src, int32_t *labels) {
    /* std::queue<int> restated as a growable ring; a vertex is pushed once per
     * successful relabel, which for unit weights is exactly once. */
    int64_t cap = n > 16 ? n : 16, head = 0, tail = 0;
    int32_t *q = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
    q[tail++] = src;
    labels[src] = 0;
    while (head < tail) {
        int32_t s = q[head++];
        for (int64_t idx = row_offsets[s]; idx < row_offsets[s + 1]; ++idx) {
            int32_t nb = col_indices[idx];
            if (labels[nb] < 0 || labels[s] + 1 < labels[nb]) {
                labels[nb] = labels[s] + 1;
                if (tail == cap) {
                    cap *= 2;
                    q = (int32_t *)realloc(q, (size_t)cap * sizeof(int32_t));
                }
                q[tail++] = nb;
            }
        }
    }
    free(q);
}

/* ------------------------------------------------------------------------- */
/* SSSP.                                                                       */
/* (1) orc_sssp_ref_preds restates sssp_problem.hxx:59-88 literally: a         */
/*     label-correcting search on a min-heap keyed by (dist[u] at push time,   */
/*     v), int distances with `int w = col_values[i]` truncation (:79), strict */
/*     `dist[v] > dist[u] + w` (:80), no stale-entry skip.  It yields the      */
/*     predecessor array the reference's test compares (test_sssp.cu:44-51)    */
/*     and (kept here, dropped there) the int distances.                       */
/* (2) orc_sssp_dist gives float distances under the GPU functor's rule        */
/*     (sssp_functor.hxx:20-29: new = labels[src] + w[e], strict <; init       */
/*     FLT_MAX / 0 at sssp_problem.hxx:44-46) by binary-heap Dijkstra in int64, */
/*     exact for integer weights while dist < 2^24.                            */
/* ------------------------------------------------------------------------- */
typedef struct { int64_t key; int32_t v; } orc_hent;
typedef struct { orc_hent *a; int64_t n, cap; } orc_heap;

static inline int hless(orc_hent x, orc_hent y) { /* std::greater<pair<int,int>> min-heap order */
    return x.key < y.key || (x.key == y.key && x.v < y.v);
}
static void hpush(orc_heap *h, int64_t key, int32_t v) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->a = (orc_hent *)realloc(h->a, (size_t)h->cap * sizeof(orc_hent)); }
    int64_t i = h->n++;
    orc_hent x = { key, v };
    while (i > 0) {
        int64_t p = (i - 1) >> 1;
        if (!hless(x, h->a[p])) break;
        h->a[i] = h->a[p]; i = p;
    }
    h->a[i] = x;
}
static orc_hent hpop(orc_heap *h) {
    orc_hent top = h->a[0], x = h->a[--h->n];
    int64_t i = 0;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && hless(h->a[c + 1], h->a[c])) ++c;
        if (!hless(h->a[c], x)) break;
        h->a[i] = h->a[c]; i = c;
    }
    if (h->n > 0) h->a[i] = x;
    return top;
}

ORC_API void orc_sssp_ref_preds(int64_t n, const int64_t *row_offsets, const int32_t *col_indices,
                                const float *col_values, int32_t src, int32_t *preds /* pre-filled -1 */,
                                int32_t *dist_out /* nullable, n entries, INT_MAX = unreached */) {
    orc_heap h = { 0, 0, 0 };
    int32_t *dist = (int32_t *)malloc((size_t)(n + 1) * sizeof(int32_t));
    for (int64_t i = 0; i <= n; ++i) dist[i] = INT32_MAX;
    hpush(&h, -1, src);              /* :69  pq.push(make_pair(-1, src)) */
    preds[src] = -1;
    dist[src] = 0;
    while (h.n) {
        int32_t u = hpop(&h).v;
        for (int64_t i = row_offsets[u]; i < row_offsets[u + 1]; ++i) {
            int32_t v = col_indices[i];
            int32_t w = (int32_t)col_values[i];           /* :79 truncation */
            if (dist[v] > dist[u] + w) {                   /* :80 */
                preds[v] = u;
                dist[v] = dist[u] + w;
                hpush(&h, dist[u], v);                     /* :83 key is dist[u], not dist[v] */
            }
        }
    }
    if (dist_out) memcpy(dist_out, dist, (size_t)n * sizeof(int32_t));
    free(dist); free(h.a);
}

ORC_API void orc_sssp_dist(int64_t n, const int64_t *row_offsets, const int32_t *col_indices,
                           const float *col_values, int32_t src, float *dist_out) {
    orc_heap h = { 0, 0, 0 };
    int64_t *dist = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) dist[i] = INT64_MAX;
    dist[src] = 0;
    hpush(&h, 0, src);
    while (h.n) {
        orc_hent t = hpop(&h);
        int32_t u = t.v;
        if (t.key > dist[u]) continue;
        for (int64_t i = row_offsets[u]; i < row_offsets[u + 1]; ++i) {
            int32_t v = col_indices[i];
            int64_t nd = dist[u] + (int64_t)col_values[i];
            if (nd < dist[v]) { dist[v] = nd; hpush(&h, nd, v); }
        }
    }
    for (int64_t i = 0; i < n; ++i) dist_out[i] = dist[i] == INT64_MAX ? FLT_MAX : (float)dist[i];
    free(dist); free(h.a);
}

/* The same rule in fp32 for weights that are not integers: new = fl(labels[src] + w[e]) exactly as the functor      */
/* computes it (sssp_functor.hxx:22), strict <.  fl(a + w) is monotone in a and >= a for w >= 0, so label-setting in  */
/* key order reaches the least fixed point -- the one every relaxation order of the GPU enactor converges to.       */
/* Non-negative floats order like their bit patterns, which are the heap keys.                                      */
ORC_API void orc_sssp_dist_f32(int64_t n, const int64_t *row_offsets, const int32_t *col_indices,
                               const float *col_values, int32_t src, float *dist_out) {
    orc_heap h = { 0, 0, 0 };
    for (int64_t i = 0; i < n; ++i) dist_out[i] = FLT_MAX;
    dist_out[src] = 0.0f;
    hpush(&h, 0, src);
    while (h.n) {
        orc_hent t = hpop(&h);
        int32_t u = t.v;
        float du = dist_out[u];
        uint32_t bits;
        memcpy(&bits, &du, sizeof bits);
        if (t.key > (int64_t)bits) continue;
        for (int64_t i = row_offsets[u]; i < row_offsets[u + 1]; ++i) {
            int32_t v = col_indices[i];
            volatile float nd = du + col_values[i];   /* (volatile: no excess precision, no contraction) */
            float ndf = nd;
            if (ndf < dist_out[v]) {
                dist_out[v] = ndf;
                memcpy(&bits, &ndf, sizeof bits);
                hpush(&h, (int64_t)bits, v);
            }
        }
    }
    free(h.a);
}

/* ------------------------------------------------------------------------- */
/* neighborhood_reduce: gunrock/src/neighborhood.hxx:47-58.  For frontier slot */
/* i holding vertex v: reduced[i] = op over u in N(v) of value[u]; `identity`  */
/* when N(v) is empty.  op: 0 = plus, 1 = min, 2 = max.  fp64 accumulation.    */
/* `values` plays Functor::get_value_to_reduce (pr_functor.hxx:27-29 for PR:   */
/* isfinite(x) ? x : 0 is applied by the caller / by orc_pr below).            */
/* ------------------------------------------------------------------------- */
ORC_API void orc_neighborhood_reduce_f64(int64_t n_front, const int32_t *frontier,
                                         const int64_t *offsets, const int32_t *indices,
                                         const double *values, int op, double identity,
                                         double *reduced, double *abs_sum /* nullable */) {
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < n_front; ++i) {
        int32_t v = frontier[i];
        int64_t b = offsets[v], e = offsets[v + 1];
        double acc = identity, as = 0.0;
        for (int64_t k = b; k < e; ++k) {
            double x = values[indices[k]];
            as += fabs(x);
            if (k == b) acc = x;
            else if (op == 0) acc += x;
            else if (op == 1) acc = x < acc ? x : acc;
            else acc = x > acc ? x : acc;
        }
        reduced[i] = acc;
        if (abs_sum) abs_sum[i] = as;
    }
}

/* PR-style driver: pr_enactor.hxx:41-77 + pr_functor.hxx:11-29 +
 * pr_problem.hxx:34-44.  Two modes:
 *   scatter = 0: reference-faithful SLOT-indexed reduced[] (neighborhood.hxx:58
 *                writes reduced[slot]; pr_functor.hxx:13 reads
 *                d_reduced_ranks[vertex]) -- SURVEY quirk 8;
 *   scatter = 1: reduced value written to reduced[vertex]
 *                (the commented-out write_reduced_value, neighborhood.hxx:60-67).
 * Sums are accumulated in fp64 and rounded to fp32 once per slot; the filter
 * arithmetic is fp32 exactly as pr_functor.hxx:11-17.
 * frontier_lens[it] receives the output length of iteration it.  Returns the
 * number of iterations run. */
ORC_API int orc_pr(int64_t n, const int64_t *offsets, const int32_t *indices, int max_iter,
                   int scatter, float *current /* n, out */, float *reduced /* n, out */,
                   int64_t *frontier_lens, const float *init /* nullable: initial ranks instead of 0.15f (pr_problem.hxx:38) */) {
    int32_t *fa = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    int32_t *fb = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    float *tmp = (float *)malloc((size_t)n * sizeof(float));
    int64_t len = n;
    for (int64_t v = 0; v < n; ++v) { fa[v] = (int32_t)v; current[v] = init ? init[v] : 0.15f; reduced[v] = 0.0f; }
    int it = 0;
    while (len > 0 && it < max_iter) {
        /* neighborhood_kernel<..., plus_t<float>, has_output=false, push=false>; CSC == CSR (graph.hxx:75-80) */
#pragma omp parallel for schedule(dynamic, 4096)
        for (int64_t i = 0; i < len; ++i) {
            int32_t v = fa[i];
            double acc = 0.0;
            for (int64_t k = offsets[v]; k < offsets[v + 1]; ++k) {
                float x = current[indices[k]];
                acc += isfinite(x) ? (double)x : 0.0;
            }
            tmp[i] = (float)acc;
        }
        for (int64_t i = 0; i < len; ++i) reduced[scatter ? fa[i] : i] = tmp[i];
        /* filter_kernel with pr_functor_t::cond_filter (stable compaction) */
        int64_t out = 0;
        for (int64_t i = 0; i < len; ++i) {
            int32_t idx = fa[i];
            float deg = (float)(offsets[idx + 1] - offsets[idx]);
            float old_value = current[idx];
            float new_value = (deg > 0) ? (0.15f + 0.85f * reduced[idx] / deg) : 0.15f;
            if (!isfinite(new_value)) new_value = 0;
            current[idx] = new_value;
            if (fabsf(new_value - old_value) > (0.001f * old_value)) fb[out++] = idx;
        }
        frontier_lens[it] = out;
        int32_t *t = fa; fa = fb; fb = t;
        len = out;
        ++it;
    }
    free(fa); free(fb); free(tmp);
    return it;
}

/* ------------------------------------------------------------------------- */
/* k-core peel: gunrock/src/kcore/kcore_problem.hxx:54-105 (the reference's CPU  */
/* validation), which the device enactor mirrors round by round                  */
/* (kcore_enactor.hxx:41-84).  For k = 1, 2, ...: repeat { every vertex with     */
/* remaining degree in (0, k) gets core number k - 1 and degree 0; stop the      */
/* rounds if nobody was peeled; every arc out of a peeled vertex lowers its head's*/
/* degree }; the first k that leaves no vertex of degree >= k ends it.  Note the  */
/* reference's quirk, kept: a vertex whose degree reaches 0 only through its      */
/* neighbours' removal is never numbered (it stays 0).  Returns the largest core. */
/* ------------------------------------------------------------------------- */
ORC_API int32_t orc_kcore(int64_t n, const int64_t *offsets, const int32_t *indices, int32_t *num_cores) {
    int64_t *deg = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int32_t *peeled = (int32_t *)malloc((size_t)(n ? n : 1) * sizeof(int32_t));
    char *remain = (char *)malloc((size_t)(n ? n : 1));
    int32_t largest = -1;
    for (int64_t v = 0; v < n; ++v) { deg[v] = offsets[v + 1] - offsets[v]; num_cores[v] = 0; }
    for (int64_t k = 1; k <= n && largest < 0; ++k) {
        int64_t num_remain = 0;
        memset(remain, 1, (size_t)n);
        for (;;) {
            int64_t np = 0;
            for (int64_t v = 0; v < n; ++v)
                if (deg[v] < k && deg[v] > 0 && remain[v]) { num_cores[v] = (int32_t)(k - 1); deg[v] = 0; peeled[np++] = (int32_t)v; }
            num_remain = 0;
            for (int64_t v = 0; v < n; ++v) { remain[v] = deg[v] >= k; num_remain += remain[v]; }
            if (np == 0) break;
            for (int64_t i = 0; i < np; ++i)
                for (int64_t e = offsets[peeled[i]]; e < offsets[peeled[i] + 1]; ++e) deg[indices[e]]--;
        }
        if (num_remain == 0) largest = (int32_t)(k - 1);
    }
    free(deg); free(peeled); free(remain);
    return largest;
}

/* ------------------------------------------------------------------------- */
/* Operator-level restatements used to check the drop-in operators one call at */
/* a time (advance.hxx:20-67, filter.hxx:11-31 with the BFS functor).          */
/* ------------------------------------------------------------------------- */

/* One push level of bfs_enactor.hxx:50-71 (advance + filter) on the CPU:
 * returns |F_next| and writes it in input order of discovery. */
ORC_API int64_t orc_bfs_push_level(const int64_t *offsets, const int32_t *indices,
                                   const int32_t *frontier, int64_t flen, int32_t iteration,
                                   int32_t *labels, int32_t *next) {
    int64_t out = 0;
    for (int64_t i = 0; i < flen; ++i) {
        int32_t v = frontier[i];
        for (int64_t k = offsets[v]; k < offsets[v + 1]; ++k) {
            int32_t u = indices[k];
            if (labels[u] == -1) { labels[u] = iteration + 1; next[out++] = u; }
        }
    }
    return out;
}

/* Sum of degrees over reached vertices: the TEPS numerator of SURVEY.md §8d. */
ORC_API int64_t orc_reached_arcs_i32(int64_t n, const int64_t *offsets, const int32_t *labels) {
    int64_t s = 0;
    for (int64_t v = 0; v < n; ++v) if (labels[v] >= 0) s += offsets[v + 1] - offsets[v];
    return s;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
