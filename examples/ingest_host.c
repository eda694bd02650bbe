/* ingest_host.c -- a plain C99 host of the C ABI (no CUDA headers, no C++): loads a MatrixMarket file with the
 * reference's load_graph semantics, keeps it in the binary CSR cache, reads it back and prints a summary.
 * With a GPU present it also uploads the graph and runs a BFS through b200_bfs_host.
 *   gcc -std=c99 -I include examples/ingest_host.c -L mini_b200 -lb200_frontier -Wl,-rpath,$PWD/mini_b200 -o build/ingest_host
 *   build/ingest_host graph.mtx [--undirected] [--bfs SRC] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b200_frontier.h"

static int fail(const char *what, int rc) {
    fprintf(stderr, "%s: %s\n", what, b200_status_string(rc));
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s file.mtx [--undirected] [--bfs SRC]\n", argv[0]);
        return 2;
    }
    int undirected = 0, bfs_src = -1;
    for (int a = 2; a < argc; ++a) {
        if (!strcmp(argv[a], "--undirected")) undirected = 1;
        else if (!strcmp(argv[a], "--bfs") && a + 1 < argc) bfs_src = atoi(argv[++a]);
    }
    b200_host_csr g, back;
    int rc = b200_mtx_load(argv[1], undirected, &g);
    if (rc != B200_OK) return fail("b200_mtx_load", rc);
    char cache[4096];
    snprintf(cache, sizeof cache, "%s.b200csr", argv[1]);
    if ((rc = b200_csr_cache_write(cache, &g)) != B200_OK) return fail("b200_csr_cache_write", rc);
    if ((rc = b200_csr_cache_read(cache, &back)) != B200_OK) return fail("b200_csr_cache_read", rc);
    int same = back.n == g.n && back.m == g.m &&
               !memcmp(back.row_offsets, g.row_offsets, sizeof(uint32_t) * (size_t)(g.n + 1)) &&
               !memcmp(back.col_indices, g.col_indices, sizeof(int32_t) * (size_t)g.m) &&
               !memcmp(back.col_values, g.col_values, sizeof(float) * (size_t)g.m);
    double wsum = 0;
    for (int64_t e = 0; e < g.m; ++e) wsum += g.col_values[e];
    printf("n=%lld m=%lld weight_sum=%.3f cache_round_trip=%s\n", (long long)g.n, (long long)g.m, wsum, same ? "ok" : "MISMATCH");

    if (bfs_src >= 0) {
        int ndev = 0;
        if (b200_device_count(&ndev) != B200_OK || ndev == 0) {
            printf("bfs skipped: no CUDA device (there is no CPU fallback)\n");
        } else {
            b200_ctx *ctx = NULL;
            b200_host_graph *hg = NULL;
            int32_t *labels = (int32_t *)malloc(sizeof(int32_t) * (size_t)g.n);
            if ((rc = b200_ctx_create(&ctx, 0, NULL)) != B200_OK) return fail("b200_ctx_create", rc);
            if ((rc = b200_host_graph_upload(ctx, g.n, g.m, g.row_offsets, g.col_indices, g.col_values, &hg)) != B200_OK)
                return fail("b200_host_graph_upload", rc);
            if ((rc = b200_bfs_host(ctx, hg, bfs_src, B200_BFS_PUSH, 0.f, 0.f, NULL, labels, NULL)) != B200_OK)
                return fail("b200_bfs_host", rc);
            int64_t reached = 0;
            int depth = 0;
            for (int64_t v = 0; v < g.n; ++v)
                if (labels[v] >= 0) {
                    ++reached;
                    if (labels[v] > depth) depth = labels[v];
                }
            printf("bfs from %d: reached=%lld depth=%d\n", bfs_src, (long long)reached, depth);
            free(labels);
            b200_host_graph_free(ctx, hg);
            b200_ctx_destroy(ctx);
        }
    }
    b200_host_csr_free(&back);
    b200_host_csr_free(&g);
    return same ? 0 : 1;
}
