// kcore_problem.hxx -- k-core decomposition problem data with the reference's member names
// (gunrock/src/kcore/kcore_problem.hxx:12-52): per-vertex core number and a working copy of the degrees.
// cpu() restates the reference's validation peel (kcore_problem.hxx:54-105) so the reference's own
// tests/kcore/test_kcore.cu runs unchanged; oracle/ref_shim.cu runs the reference's original for parity.
#pragma once
#include "problem.hxx"

namespace gunrock {
namespace kcore {

struct kcore_problem_t : problem_t {
    struct data_slice_t {
        int *d_num_cores, *d_degrees;
    };

    mem_t<int> d_num_cores, d_degrees;
    std::vector<int> num_cores, degrees;   // host mirrors: extract() / the CPU peel's working degrees
    int largest_k_core = -1;
    mem_t<data_slice_t> d_data_slice;

    kcore_problem_t() = default;

    kcore_problem_t(std::shared_ptr<graph_device_t> graph, standard_context_t &context) : problem_t(graph) {
        const int n = graph->num_nodes;
        num_cores.assign(n, 0);
        d_num_cores = mgpu::fill(0, n, context);
        d_degrees = mgpu::fill(0, n, context);
        count_degrees(context);
        degrees = from_mem(d_degrees);
        d_data_slice = publish_slice(data_slice_t{d_num_cores.data(), d_degrees.data()}, context);
    }

    // integer out-degrees (problem_t::GetDegrees yields floats for PR)
    void count_degrees(standard_context_t &context) {
        int *deg = d_degrees.data();
        const int *off = gslice->d_row_offsets.data();
        transform([=] __device__(int v) { deg[v] = off[v + 1] - off[v]; }, gslice->num_nodes, context);
    }

    void extract() { mgpu::dtoh(num_cores, d_num_cores.data(), gslice->num_nodes); }

    // Host peel, round by round like the device enactor: for k = 1, 2, ...: repeatedly give every vertex whose
    // remaining degree is in (0, k) the core number k - 1, drop it, and lower its neighbours' degrees; stop at the
    // first k that leaves no vertex of degree >= k.  Consumes `degrees`.  Returns the largest k-core.
    int cpu(std::vector<int> &validation_num_cores, std::vector<int> &row_offsets, std::vector<int> &col_indices) {
        const int n = gslice->num_nodes;
        std::vector<char> survivor(n, 1);
        std::vector<int> peeled;
        for (int k = 1; k <= n; ++k) {
            std::fill(survivor.begin(), survivor.end(), 1);
            int remain = 0;
            for (;;) {
                peeled.clear();
                for (int v = 0; v < n; ++v)
                    if (survivor[v] && degrees[v] > 0 && degrees[v] < k) {
                        validation_num_cores[v] = k - 1;
                        degrees[v] = 0;
                        peeled.push_back(v);
                    }
                remain = 0;
                for (int v = 0; v < n; ++v) remain += (survivor[v] = degrees[v] >= k);
                if (peeled.empty()) break;
                for (int v : peeled)
                    for (int e = row_offsets[v]; e < row_offsets[v + 1]; ++e) --degrees[col_indices[e]];
            }
            if (remain == 0) return k - 1;
        }
        return -1;
    }
};

}  // namespace kcore
}  // namespace gunrock
