// coloring_enactor.hxx -- hash-extrema graph colouring as a client of neighborhood_kernel with the INTEGER
// maximum / minimum reductions and of filter_kernel (gunrock/src/coloring/coloring_enactor.hxx:43-92).
//
// Every iteration: max and min of the uncoloured neighbours' hashes for EVERY vertex (the reductions run over the
// full vertex list, so their slot-indexed results line up with vertex ids), then the filter colours the local
// extrema and keeps the rest; hashes are redrawn.  Two colours per iteration, independent sets by construction.
// (The reference's call sites name five of neighborhood_kernel's six template arguments; `push` defaults to false
// here so they compile as written.)
#pragma once
#include <limits>
#include <numeric>
#include "coloring_functor.hxx"
#include "coloring_problem.hxx"
#include "enactor.hxx"
#include "filter.hxx"
#include "frontier.hxx"
#include "graph.hxx"
#include "neighborhood.hxx"
#include "test_utils.hxx"

using namespace mgpu;
using namespace gunrock::oprtr::filter;
using namespace gunrock::oprtr::neighborhood;

namespace gunrock {
namespace coloring {

struct coloring_enactor_t : enactor_t {
    std::vector<int> frontier_lengths;   // uncoloured vertices after every iteration (for tests)

    coloring_enactor_t(standard_context_t &context, int num_nodes, int num_edges) : enactor_t(context, num_nodes, num_edges) {}
    coloring_enactor_t(const coloring_enactor_t &) = delete;
    coloring_enactor_t &operator=(const coloring_enactor_t &) = delete;

    void init_frontier(std::shared_ptr<coloring_problem_t> coloring_problem, std::shared_ptr<frontier_t<int>> frontier) {
        std::vector<int> all(coloring_problem->gslice->num_nodes);
        std::iota(all.begin(), all.end(), 0);
        frontier->load(all);
        buffers[0]->load(all);
    }

    void enact(std::shared_ptr<coloring_problem_t> coloring_problem, standard_context_t &context) {
        const int n = coloring_problem->gslice->num_nodes;
        std::shared_ptr<frontier_t<int>> everyone(std::make_shared<frontier_t<int>>(context, n));
        init_frontier(coloring_problem, everyone);
        frontier_lengths.clear();
        int uncoloured = n, cur = 0;
        for (int iteration = 0; uncoloured > 0 && iteration < coloring_problem->max_iter; ++iteration) {
            neighborhood_kernel<coloring_problem_t, reduce_max_t, int, mgpu::maximum_t<int>, false>(
                coloring_problem, everyone, everyone, coloring_problem->d_reduced_max.data(), std::numeric_limits<int>::min(),
                iteration, context);
            neighborhood_kernel<coloring_problem_t, reduce_min_t, int, mgpu::minimum_t<int>, false>(
                coloring_problem, everyone, everyone, coloring_problem->d_reduced_min.data(), std::numeric_limits<int>::max(),
                iteration, context);
            uncoloured = filter_kernel<coloring_problem_t, coloring_functor_t>(coloring_problem, buffers[cur], buffers[cur ^ 1],
                                                                             iteration, context);
            frontier_lengths.push_back(uncoloured);
            cur ^= 1;
            coloring_problem->reset_hashs(context);
        }
    }
};

}  // namespace coloring
}  // namespace gunrock
