// sssp_problem.hxx -- SSSP problem data with the reference's member names
// (gunrock/src/sssp/sssp_problem.hxx:10-89): float d_labels (FLT_MAX = unreached),
// d_preds, per-iteration d_visited stamps, d_weights = the graph's d_col_values.
#pragma once
#include <limits>
#include <queue>
#include <utility>
#include "problem.hxx"

namespace gunrock {
namespace sssp {

struct sssp_problem_t : problem_t {
    mem_t<float> d_labels;
    mem_t<int> d_preds;
    mem_t<int> d_visited;
    std::vector<float> labels;
    std::vector<int> preds;
    int src = 0;

    struct data_slice_t {
        float *d_labels;
        int *d_preds;
        float *d_weights;
        int *d_visited;
        void init(mem_t<float> &_labels, mem_t<int> &_preds, mem_t<float> &_weights, mem_t<int> &_visited) {
            d_labels = _labels.data();
            d_preds = _preds.data();
            d_weights = _weights.data();
            d_visited = _visited.data();
        }
    };
    mem_t<data_slice_t> d_data_slice;
    std::vector<data_slice_t> data_slice;

    sssp_problem_t() {}
    sssp_problem_t(const sssp_problem_t &) = delete;
    sssp_problem_t &operator=(const sssp_problem_t &) = delete;

    sssp_problem_t(std::shared_ptr<graph_device_t> rhs, size_t src, standard_context_t &context)
        : problem_t(rhs), labels(rhs->num_nodes, std::numeric_limits<float>::max()), preds(rhs->num_nodes, -1),
          src((int)src), data_slice(1) {
        labels[src] = 0;
        d_labels = to_mem(labels, context);
        d_preds = to_mem(preds, context);
        d_visited = mgpu::fill(-1, rhs->num_nodes, context);
        data_slice[0].init(d_labels, d_preds, gslice->d_col_values, d_visited);
        d_data_slice = to_mem(data_slice, context);
    }

    void extract() {
        mgpu::dtoh(labels, d_labels.data(), gslice->num_nodes);
        mgpu::dtoh(preds, d_preds.data(), gslice->num_nodes);
    }

    // Host validation producing predecessors, same search as the reference
    // (sssp_problem.hxx:59-88): label-correcting on a min-heap keyed by the tail's distance,
    // integer distances (weights truncated to int), strict improvement.
    void cpu(std::vector<int> &validation_preds, std::vector<int> &row_offsets, std::vector<int> &col_indices,
             std::vector<float> &col_values) {
        typedef std::pair<int, int> entry_t;   // (key, vertex)
        std::priority_queue<entry_t, std::vector<entry_t>, std::greater<entry_t>> heap;
        std::vector<int> dist(row_offsets.size(), std::numeric_limits<int>::max());
        dist[src] = 0;
        validation_preds[src] = -1;
        heap.push(entry_t(-1, src));
        while (!heap.empty()) {
            const int u = heap.top().second;
            heap.pop();
            for (int k = row_offsets[u]; k < row_offsets[u + 1]; ++k) {
                const int v = col_indices[k];
                const int through_u = dist[u] + (int)col_values[k];
                if (through_u < dist[v]) {
                    dist[v] = through_u;
                    validation_preds[v] = u;
                    heap.push(entry_t(dist[u], v));
                }
            }
        }
    }

    // Host validation of the DISTANCES (what the device computes deterministically; the
    // reference's test only compares the racy preds).  Dijkstra with lazy deletion.
    void cpu_distances(std::vector<float> &validation_labels, std::vector<int> &row_offsets,
                       std::vector<int> &col_indices, std::vector<float> &col_values) {
        typedef std::pair<float, int> entry_t;
        std::priority_queue<entry_t, std::vector<entry_t>, std::greater<entry_t>> heap;
        validation_labels.assign(row_offsets.size() - 1, std::numeric_limits<float>::max());
        validation_labels[src] = 0;
        heap.push(entry_t(0.0f, src));
        while (!heap.empty()) {
            const entry_t top = heap.top();
            heap.pop();
            if (top.first > validation_labels[top.second]) continue;
            for (int k = row_offsets[top.second]; k < row_offsets[top.second + 1]; ++k) {
                const float nd = top.first + col_values[k];
                if (nd < validation_labels[col_indices[k]]) {
                    validation_labels[col_indices[k]] = nd;
                    heap.push(entry_t(nd, col_indices[k]));
                }
            }
        }
    }
};

}  // namespace sssp
}  // namespace gunrock
