// bfs_problem.hxx -- BFS problem data (labels / preds) with the reference's member names
// (gunrock/src/bfs/bfs_problem.hxx:9-73): d_labels = -1 for unvisited else depth, src depth 0.
#pragma once
#include <deque>
#include "problem.hxx"

namespace gunrock {
namespace bfs {

struct bfs_problem_t : problem_t {
    mem_t<int> d_labels;
    mem_t<int> d_preds;
    std::vector<int> labels;
    std::vector<int> preds;
    int src = 0;

    struct data_slice_t {
        int *d_labels;
        int *d_preds;
        void init(mem_t<int> &_labels, mem_t<int> &_preds) {
            d_labels = _labels.data();
            d_preds = _preds.data();
        }
    };
    mem_t<data_slice_t> d_data_slice;
    std::vector<data_slice_t> data_slice;

    bfs_problem_t() {}
    bfs_problem_t(const bfs_problem_t &) = delete;
    bfs_problem_t &operator=(const bfs_problem_t &) = delete;

    bfs_problem_t(std::shared_ptr<graph_device_t> rhs, size_t src, standard_context_t &context)
        : problem_t(rhs), labels(rhs->num_nodes, -1), preds(rhs->num_nodes, -1), src((int)src), data_slice(1) {
        labels[src] = 0;
        d_labels = to_mem(labels, context);
        d_preds = to_mem(preds, context);
        data_slice[0].init(d_labels, d_preds);
        d_data_slice = to_mem(data_slice, context);
    }

    void extract() { mgpu::dtoh(labels, d_labels.data(), gslice->num_nodes); }

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
