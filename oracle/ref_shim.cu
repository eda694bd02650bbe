// ref_shim.cu -- C entry points around the UNMODIFIED reference CPU code.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  Compiled by
// oracle/Makefile into oracle/_ref/libref_cpu.so with the reference headers
// included from where they lie under /root/reference (never copied).  It calls
//   gunrock::bfs::bfs_problem_t::cpu    gunrock/src/bfs/bfs_problem.hxx:52-72
//   gunrock::sssp::sssp_problem_t::cpu  gunrock/src/sssp/sssp_problem.hxx:59-88
//   gunrock::load_graph                 gunrock/src/graph.hxx:96-223
// on default-constructed problems (no CUDA call is made, so it runs on a box
// without a GPU).  Used to pin oracle.c and as bench.py's `--impl reference`
// CPU arm (cpu_baseline.kind = "reference") for graphs the reference's int32
// CSR can hold (< 2^31 arcs).
#include "bfs/bfs_problem.hxx"
#include "sssp/sssp_problem.hxx"
#include <cstring>
#include <chrono>

using namespace gunrock;

extern "C" {

// labels: n ints, filled with -1 here exactly as tests/bfs/test_bfs.cu:44 does.
// Returns seconds spent inside cpu() (std::chrono::steady_clock).
double ref_bfs_cpu(int n, int m, const int *offsets, const int *indices, int src, int *labels) {
    std::vector<int> off(offsets, offsets + n + 1), ind(indices, indices + m);
    std::vector<int> val(n, -1);
    bfs::bfs_problem_t p;
    p.src = src;
    auto t0 = std::chrono::steady_clock::now();
    p.cpu(val, off, ind);
    auto t1 = std::chrono::steady_clock::now();
    std::memcpy(labels, val.data(), sizeof(int) * (size_t)n);
    return std::chrono::duration<double>(t1 - t0).count();
}

// preds: n ints (tests/sssp/test_sssp.cu:44 initialises them to -1)
double ref_sssp_cpu(int n, int m, const int *offsets, const int *indices, const float *weights,
                    int src, int *preds) {
    std::vector<int> off(offsets, offsets + n + 1), ind(indices, indices + m);
    std::vector<float> w(weights, weights + m);
    std::vector<int> val(n, -1);
    sssp::sssp_problem_t p;
    p.src = src;
    auto t0 = std::chrono::steady_clock::now();
    p.cpu(val, off, ind, w);
    auto t1 = std::chrono::steady_clock::now();
    std::memcpy(preds, val.data(), sizeof(int) * (size_t)n);
    return std::chrono::duration<double>(t1 - t0).count();
}

// (kcore: gunrock/src/kcore/kcore_problem.hxx does not compile -- its constructor hands a mem_t<int> to
// problem_t::GetDegrees(mem_t<float>&), kcore_problem.hxx:44 -- so its cpu() cannot be built unmodified;
// oracle.c restates it, unpinned.)

// load_graph + copy-out.  First call with offsets == NULL to get sizes.
// NOTE: the reference's sort comparator is not a strict weak order
// (graph.hxx:139-157); only feed it files without duplicate edges.
int ref_load_graph(const char *file, int undirected, int *n_out, int *m_out,
                   int *offsets, int *indices, float *weights, int *csc_equals_csr) {
    std::shared_ptr<graph_t> g = load_graph(file, undirected != 0, false);
    if (!g) return 1;
    *n_out = g->num_nodes;
    *m_out = g->num_edges;
    if (offsets) {
        std::memcpy(offsets, g->csr->offsets.data(), sizeof(int) * (size_t)(g->num_nodes + 1));
        std::memcpy(indices, g->csr->indices.data(), sizeof(int) * (size_t)g->num_edges);
        std::memcpy(weights, g->csr->edge_weights.data(), sizeof(float) * (size_t)g->num_edges);
    }
    if (csc_equals_csr)
        *csc_equals_csr = (g->csc->offsets == g->csr->offsets && g->csc->indices == g->csr->indices) ? 1 : 0;
    return 0;
}

}  // extern "C"
