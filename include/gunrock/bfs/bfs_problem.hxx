// bfs_problem.hxx -- BFS problem data (labels / preds) with the reference's member names
// (gunrock/src/bfs/bfs_problem.hxx:9-73): d_labels = -1 for unvisited else depth, src depth 0.
#pragma once
#include <deque>
#include "problem.hxx"

namespace gunrock {
namespace bfs {

struct bfs_problem_t : problem_t {
    // what bfs_functor_t reads and writes, as raw device pointers
    // (d_visited is an addition: one bit per vertex, set together with the label.  The reference's functor tests and
    // claims the 4-byte label itself, bfs_functor.hxx:26-33 -- 16 MB of random probes at scale 22 against 512 KB here,
    // which is what lets the probes of advance_forward_kernel live in L1.)
    struct data_slice_t {
        int *d_labels, *d_preds;
        unsigned int *d_visited;
    };

    int src = 0;
    std::vector<int> labels, preds;        // host copies (extract() refreshes labels)
    mem_t<int> d_labels, d_preds;
    mem_t<unsigned int> d_visited;
    mem_t<data_slice_t> d_data_slice;

    bfs_problem_t() = default;

    bfs_problem_t(std::shared_ptr<graph_device_t> graph, size_t source, standard_context_t &context)
        : problem_t(graph), src((int)source), labels(graph->num_nodes, -1), preds(graph->num_nodes, -1) {
        labels[source] = 0;                // the source is at depth 0, everything else unvisited
        d_labels = to_mem(labels, context);
        d_preds = to_mem(preds, context);
        std::vector<unsigned int> bits(((size_t)graph->num_nodes + 31) / 32 + 1, 0u);
        bits[source >> 5] |= 1u << (source & 31);
        d_visited = to_mem(bits, context);
        d_data_slice = publish_slice(data_slice_t{d_labels.data(), d_preds.data(), d_visited.data()}, context);
    }

    void extract() { mgpu::dtoh(labels, d_labels.data(), gslice->num_nodes); }

    // The same data as the engine's POD (b200_frontier.h), for calling the C ABI operators on this problem.
    b200_problem engine_view(uint32_t *d_visited_bitmap = nullptr) {
        b200_problem p = {};
        p.kind = B200_PROBLEM_BFS;
        p.labels = d_labels.data();
        p.preds = d_preds.data();
        p.visited_bitmap = d_visited_bitmap ? d_visited_bitmap : d_visited.data();
        return p;
    }

    // Host validation: level-synchronous BFS over a FIFO, same result as the reference's
    // queue BFS (bfs_problem.hxx:52-72).  validation_labels must come in as all -1.
    void cpu(std::vector<int> &validation_labels, std::vector<int> &row_offsets, std::vector<int> &col_indices) {
        std::deque<int> fifo(1, src);
        validation_labels[src] = 0;
        while (!fifo.empty()) {
            const int u = fifo.front();
            fifo.pop_front();
            const int next = validation_labels[u] + 1;
            for (int k = row_offsets[u]; k < row_offsets[u + 1]; ++k) {
                const int v = col_indices[k];
                if (validation_labels[v] < 0 || next < validation_labels[v]) {
                    validation_labels[v] = next;
                    fifo.push_back(v);
                }
            }
        }
    }
};

}  // namespace bfs
}  // namespace gunrock
