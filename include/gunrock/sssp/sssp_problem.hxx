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
    // what sssp_functor_t reads and writes, as raw device pointers (member ORDER as in the reference: labels, preds,
    // weights, visited)
    struct data_slice_t {
        float *d_labels;
        int *d_preds;
        float *d_weights;
        int *d_visited;
    };

    int src = 0;
    std::vector<float> labels;             // host copies, refreshed by extract()
    std::vector<int> preds;
    mem_t<float> d_labels;
    mem_t<int> d_preds, d_visited;
    mem_t<data_slice_t> d_data_slice;

    sssp_problem_t() = default;

    sssp_problem_t(std::shared_ptr<graph_device_t> graph, size_t source, standard_context_t &context)
        : problem_t(graph), src((int)source), labels(graph->num_nodes, std::numeric_limits<float>::max()),
          preds(graph->num_nodes, -1) {
        labels[source] = 0;                // unreached = FLT_MAX, the source is at distance 0
        d_labels = to_mem(labels, context);
        d_preds = to_mem(preds, context);
        d_visited = mgpu::fill(-1, graph->num_nodes, context);   // per-iteration dedupe stamps
        d_data_slice = publish_slice(
            data_slice_t{d_labels.data(), d_preds.data(), gslice->d_col_values.data(), d_visited.data()}, context);
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
