// ref_gpu_driver.cu -- runs the UNMODIFIED reference GPU primitives on a CSR handed over in a file.
//
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/oracle.c header).  Compiled by oracle/Makefile for
// sm_100 into oracle/_ref/ref_gpu_{bfs,sssp,pr} with the reference headers included from where they lie under
// /root/reference (never copied).  A separate PROCESS on purpose: the reference exits the process on a
// frontier overflow (frontier.hxx:53-59) and prints to stdout from its enactors.  It calls
//   gunrock::bfs::bfs_enactor_t::enact_pushpull    gunrock/src/bfs/bfs_enactor.hxx:41-117   (as tests/bfs/test_bfs.cu:26-42)
//   gunrock::sssp::sssp_enactor_t::enact           gunrock/src/sssp/sssp_enactor.hxx:40-72  (as tests/sssp/test_sssp.cu:31-43)
//   gunrock::pr::pr_enactor_t::enact               gunrock/src/pr/pr_enactor.hxx:41-79      (as tests/pr/test_pr.cu:26-40)
// and is used (a) to produce the golden vectors that pin the neighbourhood-reduce / PR part of the oracle
// (tests/golden/make_golden_gpu.py, run on the GPU box) and (b) as the "reference GPU" baseline column of
// bench.py (wall clock around enact(), exactly what the reference's tests print as "elapsed time").
//
//   ref_gpu_<prim> <bfs|sssp|pr> <in.bin> <out.bin> [--src=S] [--alpha=A] [--max_iter=K] [--runs=R]
//                  [--queue-sizing=Q] [--values=<f32 file: initial current_ranks, pr only>]
// in.bin : int64 n, int64 m, int32 has_weights, int32 undirected, int32 offsets[n+1], int32 indices[m], f32 weights[m]
// out.bin: bfs: int32 labels[n] | sssp: f32 labels[n], int32 preds[n] | pr: f32 current[n], f32 reduced[n]
// Last stdout line: one JSON object with the per-run wall-clock seconds.
// (advance.hxx / filter.hxx carry no include guard, so one primitive per translation unit: the Makefile builds
// ref_gpu_bfs / ref_gpu_sssp / ref_gpu_pr from this file with -DREF_PRIM_BFS / _SSSP / _PR, include order as in
// the reference's own tests.)
#if defined(REF_PRIM_BFS)
#include "bfs/bfs_enactor.hxx"
#define REF_PRIM "bfs"
#elif defined(REF_PRIM_SSSP)
#include "sssp/sssp_enactor.hxx"
#define REF_PRIM "sssp"
#elif defined(REF_PRIM_PR)
#include "pr/pr_enactor.hxx"
#define REF_PRIM "pr"
#else
#error "define REF_PRIM_BFS, REF_PRIM_SSSP or REF_PRIM_PR"
#endif
#include "test_utils.hxx"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

using namespace gunrock;

static bool read_all(FILE *f, void *p, size_t bytes) { return bytes == 0 || std::fread(p, 1, bytes, f) == bytes; }

int main(int argc, char **argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <bfs|sssp|pr> <in.bin> <out.bin> [--src=S] [--alpha=A] [--max_iter=K] [--runs=R]\n", argv[0]);
        return 2;
    }
    const std::string prim = argv[1];
    if (prim != REF_PRIM) {
        std::fprintf(stderr, "this binary runs %s only\n", REF_PRIM);
        return 2;
    }
    CommandLineArgs args(argc, argv);
    int src = 0, max_iter = 10, runs = 1;
    float queue_sizing = 1.0f;
    std::string values_file;
    args.GetCmdLineArgument("src", src);
    args.GetCmdLineArgument("max_iter", max_iter);
    args.GetCmdLineArgument("runs", runs);
    args.GetCmdLineArgument("queue-sizing", queue_sizing);
    args.GetCmdLineArgument("values", values_file);

    FILE *f = std::fopen(argv[2], "rb");
    if (!f) { std::perror(argv[2]); return 2; }
    long long n = 0, m = 0;
    int has_w = 0, undirected = 1;
    if (!read_all(f, &n, 8) || !read_all(f, &m, 8) || !read_all(f, &has_w, 4) || !read_all(f, &undirected, 4) ||
        n < 1 || m < 0 || m >= (1ll << 31) || n >= (1ll << 31)) {
        std::fprintf(stderr, "bad header\n");
        return 2;
    }
    std::shared_ptr<graph_t> graph(std::make_shared<graph_t>());
    graph->undirected = undirected != 0;
    graph->num_nodes = (int)n;
    graph->num_edges = (int)m;
    graph->csr = std::make_shared<csr_t>();
    csr_t &c = *graph->csr;
    c.num_nodes = (int)n;
    c.num_edges = (int)m;
    c.offsets.resize(n + 1);
    c.indices.resize(m);
    c.edge_weights.assign(m, 1.0f);
    if (!read_all(f, c.offsets.data(), 4 * (size_t)(n + 1)) || !read_all(f, c.indices.data(), 4 * (size_t)m) ||
        (has_w && !read_all(f, c.edge_weights.data(), 4 * (size_t)m))) {
        std::fprintf(stderr, "short file\n");
        return 2;
    }
    std::fclose(f);
    c.sources.resize(m);   // what load_graph leaves there (graph.hxx:159-172): the source of every arc
    for (long long v = 0; v < n; ++v)
        for (int e = c.offsets[v]; e < c.offsets[v + 1]; ++e) c.sources[e] = (int)v;
    graph->csc = graph->csr;   // symmetric inputs only: graph_to_device aliases CSC to CSR (graph.hxx:75-80)

    standard_context_t context;
    std::shared_ptr<graph_device_t> d_graph(std::make_shared<graph_device_t>());
    graph_to_device(d_graph, graph, context);

    std::vector<double> secs;
    FILE *o = std::fopen(argv[3], "wb");
    if (!o) { std::perror(argv[3]); return 2; }
    bool ok = true;
    auto put = [&](const void *p, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, o) == bytes); };

#if defined(REF_PRIM_BFS)
    {
        float alpha = 1.0f / d_graph->num_nodes;   // tests/bfs/test_bfs.cu:30
        args.GetCmdLineArgument("alpha", alpha);
        std::vector<int> labels;
        for (int r = 0; r < runs; ++r) {
            std::shared_ptr<bfs::bfs_problem_t> p(std::make_shared<bfs::bfs_problem_t>(d_graph, src, context));
            std::shared_ptr<bfs::bfs_enactor_t> e(std::make_shared<bfs::bfs_enactor_t>(context, d_graph->num_nodes, d_graph->num_edges));
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            e->enact_pushpull(p, alpha, context);
            cudaDeviceSynchronize();
            secs.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            p->extract();
            labels = p->labels;
        }
        put(labels.data(), 4 * (size_t)n);
    }
#elif defined(REF_PRIM_SSSP)
    {
        std::vector<float> labels;
        std::vector<int> preds;
        for (int r = 0; r < runs; ++r) {
            std::shared_ptr<sssp::sssp_problem_t> p(std::make_shared<sssp::sssp_problem_t>(d_graph, src, context));
            std::shared_ptr<sssp::sssp_enactor_t> e(
                std::make_shared<sssp::sssp_enactor_t>(context, d_graph->num_nodes, d_graph->num_edges, queue_sizing));
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            e->enact(p, context);
            cudaDeviceSynchronize();
            secs.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            p->extract();
            labels = p->labels;
            preds = p->preds;
        }
        put(labels.data(), 4 * (size_t)n);
        put(preds.data(), 4 * (size_t)n);
    }
#else
    {
        std::vector<float> cur(n), red(n), init;
        if (!values_file.empty()) {
            FILE *vf = std::fopen(values_file.c_str(), "rb");
            init.resize(n);
            if (!vf || !read_all(vf, init.data(), 4 * (size_t)n)) { std::fprintf(stderr, "bad --values file\n"); return 2; }
            std::fclose(vf);
        }
        for (int r = 0; r < runs; ++r) {
            std::shared_ptr<pr::pr_problem_t> p(std::make_shared<pr::pr_problem_t>(d_graph, max_iter, context));
            if (!init.empty()) mgpu::htod(p->d_current_ranks.data(), init.data(), (size_t)n);
            std::shared_ptr<pr::pr_enactor_t> e(std::make_shared<pr::pr_enactor_t>(context, d_graph->num_nodes, d_graph->num_edges));
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            e->enact(p, context);
            cudaDeviceSynchronize();
            secs.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            mgpu::dtoh(cur.data(), p->d_current_ranks.data(), (size_t)n);
            mgpu::dtoh(red.data(), p->d_reduced_ranks.data(), (size_t)n);
        }
        put(cur.data(), 4 * (size_t)n);
        put(red.data(), 4 * (size_t)n);
    }
#endif
    ok = (std::fclose(o) == 0) && ok;
    const cudaError_t ce = cudaDeviceSynchronize();
    std::printf("\n{\"prim\": \"%s\", \"n\": %lld, \"m\": %lld, \"cuda_error\": %d, \"write_ok\": %s, \"elapsed_s\": [", prim.c_str(), n, m,
                (int)ce, ok ? "true" : "false");
    for (size_t i = 0; i < secs.size(); ++i) std::printf("%s%.9f", i ? ", " : "", secs[i]);
    std::printf("]}\n");
    return (ok && ce == cudaSuccess) ? 0 : 1;
}
