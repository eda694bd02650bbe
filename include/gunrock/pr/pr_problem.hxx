// pr_problem.hxx -- PageRank-style problem data with the reference's member names
// (gunrock/src/pr/pr_problem.hxx:8-49): ranks start at 0.15, degrees as floats.
#pragma once
#include "problem.hxx"

namespace gunrock {
namespace pr {

struct pr_problem_t : problem_t {
    // what pr_functor_t reads and writes, as raw device pointers
    struct data_slice_t {
        float *d_current_ranks, *d_reduced_ranks, *d_degrees;
    };

    static constexpr float kInitialRank = 0.15f;   // pr_problem.hxx:38
    int max_iter = 0;
    mem_t<float> d_current_ranks, d_reduced_ranks, d_degrees;
    mem_t<data_slice_t> d_data_slice;

    pr_problem_t() = default;

    pr_problem_t(std::shared_ptr<graph_device_t> graph, int iterations, standard_context_t &context)
        : problem_t(graph), max_iter(iterations) {
        const int n = graph->num_nodes;
        d_current_ranks = mgpu::fill(kInitialRank, n, context);
        d_reduced_ranks = mgpu::fill(0.0f, n, context);
        d_degrees = mgpu::fill(0.0f, n, context);
        GetDegrees(d_degrees, context);
        d_data_slice = publish_slice(data_slice_t{d_current_ranks.data(), d_reduced_ranks.data(), d_degrees.data()}, context);
    }

    void extract() {}   // (the reference's is empty too: test_pr.cu never reads the ranks back)
};

}  // namespace pr
}  // namespace gunrock
