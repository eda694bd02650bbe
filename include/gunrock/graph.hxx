// graph.hxx -- host CSR, host graph, device graph and the MatrixMarket loader.
// Same type and member names as the reference (gunrock/src/graph.hxx:19-83) so
// primitives written against it compile unchanged; the bodies are new.
//
// Loader semantics kept from graph.hxx:96-223: 1-based `i j [w]` lines, a line
// becomes the arc j -> i (rows are built over the second column), missing weights
// are 1.0 (or rand()%64 with _random_edge_value), `_undir` appends every reverse,
// duplicates and self loops are kept, arcs end up ordered by (row, column).
// Differences, all deliberate:
//   * the ordering uses a strict-weak comparator + stable sort (the reference's
//     comparator returns true on equal keys, graph.hxx:139-157: UB with duplicates);
//   * graph_t::undirected is set to _undir (the reference stores !_undir and then
//     aliases CSC to CSR in every case, graph.hxx:176-177,215-222).  The net effect
//     is kept: CSC is a copy of CSR.  Define B200_TRUE_CSC to get the real
//     transpose for directed inputs instead.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include "mgpu_compat.hxx"

using namespace mgpu;

namespace gunrock {

struct csr_t {
    int num_nodes;
    int num_edges;
    std::vector<int> offsets;
    std::vector<int> indices;
    std::vector<float> edge_weights;
    std::vector<int> sources;
};

struct graph_t {
    bool undirected;
    int num_nodes;
    int num_edges;
    std::shared_ptr<csr_t> csr;
    std::shared_ptr<csr_t> csc;
};

struct graph_device_t {
    int num_nodes = 0;
    int num_edges = 0;
    mem_t<int> d_row_offsets;
    mem_t<int> d_col_indices;
    mem_t<float> d_col_values;
    mem_t<int> d_col_offsets;
    mem_t<int> d_row_indices;
    mem_t<float> d_row_values;
    mem_t<int> d_csr_srcs;
    mem_t<int> d_csc_srcs;
    // exclusive scan of the current frontier's degrees (filled by the operators; the
    // engine keeps its own copy in the workspace, this one mirrors it for clients)
    mem_t<int> d_scanned_row_offsets;

    // Engine view of this graph (offsets reinterpreted as unsigned, see b200_frontier.h).
    b200_graph view() const {
        b200_graph g;
        g.n = num_nodes;
        g.m = num_edges;
        g.row_offsets = reinterpret_cast<const uint32_t *>(d_row_offsets.data());
        g.col_indices = d_col_indices.data();
        g.col_values = d_col_values.data();
        g.col_offsets = reinterpret_cast<const uint32_t *>(d_col_offsets.data());
        g.row_indices = d_row_indices.data();
        g.row_values = d_row_values.data();
        g.no_in_arc_bitmap = nullptr;   // the engine derives it per traversal
        g.first_in_neighbor = nullptr;
        g.hot_ids = nullptr;
        g.hot_indices = nullptr;
        g.hot_count = 0;
        return g;
    }
};

// Host graph -> device graph (graph.hxx:60-83).
inline void graph_to_device(std::shared_ptr<graph_device_t> d_graph, std::shared_ptr<graph_t> graph,
                            standard_context_t &context) {
    const csr_t &out = *graph->csr;
    const csr_t &in = graph->undirected ? *graph->csr : *graph->csc;
    d_graph->num_nodes = graph->num_nodes;
    d_graph->num_edges = graph->num_edges;
    d_graph->d_row_offsets = to_mem(out.offsets, context);
    d_graph->d_col_indices = to_mem(out.indices, context);
    d_graph->d_col_values = to_mem(out.edge_weights, context);
    d_graph->d_csr_srcs = to_mem(out.sources, context);
    d_graph->d_col_offsets = to_mem(in.offsets, context);
    d_graph->d_row_indices = to_mem(in.indices, context);
    d_graph->d_row_values = to_mem(in.edge_weights, context);
    d_graph->d_csc_srcs = to_mem(in.sources, context);
    d_graph->d_scanned_row_offsets = mem_t<int>(graph->num_nodes, context);
    b200_ctx_reserve(context.engine(), std::max(graph->num_nodes, graph->num_edges));
}

inline void display_csr(std::shared_ptr<csr_t> csr) {
    std::cout << "offsets: \n";
    for (int x : csr->offsets) std::cout << x << ' ';
    std::cout << "\nindices: \n";
    for (int x : csr->indices) std::cout << x << ' ';
    std::cout << std::endl;
}

namespace detail {
struct arc_t {
    int row, col;
    float w;
};
// rows = arc.row; arcs ordered by (row, col), ties keep input order.
inline std::shared_ptr<csr_t> csr_from_arcs(int n, std::vector<arc_t> arcs) {
    std::stable_sort(arcs.begin(), arcs.end(), [](const arc_t &a, const arc_t &b) {
        return a.row != b.row ? a.row < b.row : a.col < b.col;
    });
    auto csr = std::make_shared<csr_t>();
    const int m = (int)arcs.size();
    csr->num_nodes = n;
    csr->num_edges = m;
    csr->offsets.assign(n + 1, 0);
    csr->indices.resize(m);
    csr->edge_weights.resize(m);
    csr->sources.resize(m);
    for (const arc_t &a : arcs) csr->offsets[a.row + 1]++;
    std::partial_sum(csr->offsets.begin(), csr->offsets.end(), csr->offsets.begin());
    for (int e = 0; e < m; ++e) {
        csr->indices[e] = arcs[e].col;
        csr->edge_weights[e] = arcs[e].w;
        csr->sources[e] = arcs[e].row;
    }
    return csr;
}
}  // namespace detail

// Build a graph_t straight from (row, col, weight) arcs: the entry point synthetic
// generators use instead of pushing RMAT through a .mtx file.
inline std::shared_ptr<graph_t> graph_from_arcs(int num_vertices, const std::vector<int> &rows,
                                                const std::vector<int> &cols, const std::vector<float> &weights,
                                                bool symmetric) {
    std::vector<detail::arc_t> arcs(rows.size());
    for (size_t e = 0; e < rows.size(); ++e) arcs[e] = {rows[e], cols[e], weights.empty() ? 1.0f : weights[e]};
    auto csr = detail::csr_from_arcs(num_vertices, arcs);
    std::shared_ptr<csr_t> csc;
    if (symmetric) {
        csc = std::make_shared<csr_t>(*csr);
    } else {
        for (auto &a : arcs) std::swap(a.row, a.col);
        csc = detail::csr_from_arcs(num_vertices, arcs);
    }
    return std::shared_ptr<graph_t>(new graph_t{symmetric, num_vertices, csr->num_edges, csr, csc});
}

inline std::shared_ptr<graph_t> load_graph(const char *_name, bool _undir = false, bool _random_edge_value = false) {
    std::ifstream in(_name);
    if (!in) return nullptr;
    std::string line;
    while (std::getline(in, line) && !line.empty() && line[0] == '%') {}
    int height = 0, width = 0, declared = 0;
    if (std::sscanf(line.c_str(), "%d %d %d", &height, &width, &declared) != 3) {
        std::printf("Error reading %s\n", _name);
        std::exit(0);
    }
    std::vector<detail::arc_t> arcs;
    arcs.reserve((size_t)declared * (_undir ? 2 : 1));
    std::vector<detail::arc_t> reversed;
    for (int e = 0; e < declared; ++e) {
        int i = 0, j = 0;
        float w = 0;
        int got = 0;
        if (!std::getline(in, line) || (got = std::sscanf(line.c_str(), "%d %d %f", &i, &j, &w)) < 2) {
            std::printf("Error reading edge lists %s\n", _name);
            std::exit(0);
        }
        if (got == 2) w = _random_edge_value ? (float)(std::rand() % 64) : 1.0f;
        arcs.push_back({j - 1, i - 1, w});                  // line "i j" is the arc j -> i
        if (_undir) reversed.push_back({i - 1, j - 1, w});
    }
    arcs.insert(arcs.end(), reversed.begin(), reversed.end());
    auto csr = detail::csr_from_arcs(height, arcs);
    std::shared_ptr<csr_t> csc = std::make_shared<csr_t>(*csr);
#ifdef B200_TRUE_CSC
    if (!_undir) {
        for (auto &a : arcs) std::swap(a.row, a.col);
        csc = detail::csr_from_arcs(height, arcs);
    }
#endif
    return std::shared_ptr<graph_t>(new graph_t{_undir, height, csr->num_edges, csr, csc});
}

}  // namespace gunrock
