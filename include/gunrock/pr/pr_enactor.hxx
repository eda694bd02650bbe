// pr_enactor.hxx -- PageRank-style iteration over a shrinking frontier: pull-sum the
// neighbours' ranks with neighborhood_kernel, then filter_kernel updates the ranks and
// keeps the vertices that have not converged (gunrock/src/pr/pr_enactor.hxx:41-79).
// The reduced sums are indexed by FRONTIER SLOT exactly as in the reference (SURVEY
// quirk 8); enact_scatter() is the vertex-indexed variant via write_reduced_value.
#pragma once
#include <numeric>
#include "enactor.hxx"
#include "filter.hxx"
#include "frontier.hxx"
#include "graph.hxx"
#include "neighborhood.hxx"
#include "pr_functor.hxx"
#include "pr_problem.hxx"
#include "test_utils.hxx"

using namespace mgpu;
using namespace gunrock::oprtr::filter;
using namespace gunrock::oprtr::neighborhood;

namespace gunrock {
namespace pr {

struct pr_enactor_t : enactor_t {
    std::vector<int> frontier_lengths;   // output length of every iteration (for tests)

    pr_enactor_t(standard_context_t &context, int num_nodes, int num_edges) : enactor_t(context, num_nodes, num_edges) {}
    pr_enactor_t(const pr_enactor_t &) = delete;
    pr_enactor_t &operator=(const pr_enactor_t &) = delete;

    void init_frontier(std::shared_ptr<pr_problem_t> pr_problem) {
        std::vector<int> all(pr_problem->gslice->num_nodes);
        std::iota(all.begin(), all.end(), 0);
        buffers[0]->load(all);
    }

    template <bool scatter>
    void run(std::shared_ptr<pr_problem_t> pr_problem, standard_context_t &context) {
        init_frontier(pr_problem);
        frontier_lengths.clear();
        int frontier_length = pr_problem->gslice->num_nodes, cur = 0, iteration = 0;
        // scatter: sums are produced per slot in a scratch array, then write_reduced_value moves
        // them to d_reduced_ranks[vertex] (in-place scatter would race between slots)
        mem_t<float> slot_sums;
        if (scatter) slot_sums = mem_t<float>(pr_problem->gslice->num_nodes, context);
        float *reduced = scatter ? slot_sums.data() : pr_problem->d_reduced_ranks.data();
        std::shared_ptr<frontier_t<int>> unused(std::make_shared<frontier_t<int>>(context, 1));
        while (frontier_length > 0 && iteration < pr_problem->max_iter) {
            neighborhood_kernel<pr_problem_t, pr_functor_t, float, mgpu::plus_t<float>, false, false, scatter>(
                pr_problem, buffers[cur], unused, reduced, 0.0f, iteration, context);
            frontier_length = filter_kernel<pr_problem_t, pr_functor_t>(pr_problem, buffers[cur], buffers[cur ^ 1], iteration, context);
            std::cout << "finished iteration:" << iteration << " output length: " << frontier_length << std::endl;
            frontier_lengths.push_back(frontier_length);
            ++iteration;
            cur ^= 1;
        }
        display_device_data(pr_problem->d_current_ranks.data(), std::min(10, pr_problem->gslice->num_nodes));
    }

    void enact(std::shared_ptr<pr_problem_t> pr_problem, standard_context_t &context) { run<false>(pr_problem, context); }
    void enact_scatter(std::shared_ptr<pr_problem_t> pr_problem, standard_context_t &context) { run<true>(pr_problem, context); }
};

}  // namespace pr
}  // namespace gunrock
