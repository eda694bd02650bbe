// kcore_enactor.hxx -- k-core decomposition as a client of filter_kernel and of
// advance_forward_kernel<..., idempotence = false, has_output = false> (the one primitive of the reference that
// uses the output-less advance; gunrock/src/kcore/kcore_enactor.hxx:41-84).
//
// For k = 1, 2, ...: peel rounds over the full vertex list -- filter out the vertices whose remaining degree
// dropped below k (that stamps their core number), advance over them to lower their neighbours' degrees, count
// who is left -- until a round peels nobody; the first k that leaves nobody is one past the largest core.
#pragma once
#include <numeric>
#include "advance.hxx"
#include "enactor.hxx"
#include "filter.hxx"
#include "frontier.hxx"
#include "graph.hxx"
#include "kcore_functor.hxx"
#include "kcore_problem.hxx"
#include "test_utils.hxx"

using namespace mgpu;
using namespace gunrock::oprtr::advance;
using namespace gunrock::oprtr::filter;

namespace gunrock {
namespace kcore {

struct kcore_enactor_t : enactor_t {
    kcore_enactor_t(standard_context_t &context, int num_nodes, int num_edges) : enactor_t(context, num_nodes, num_edges) {}
    kcore_enactor_t(const kcore_enactor_t &) = delete;
    kcore_enactor_t &operator=(const kcore_enactor_t &) = delete;

    void init_frontier(std::shared_ptr<kcore_problem_t> kcore_problem) {
        std::vector<int> all(kcore_problem->gslice->num_nodes);
        std::iota(all.begin(), all.end(), 0);
        buffers[0]->load(all);
    }

    void enact(std::shared_ptr<kcore_problem_t> kcore_problem, standard_context_t &context) {
        const int num_nodes = kcore_problem->gslice->num_nodes;
        std::shared_ptr<frontier_t<int>> &everyone = buffers[0], &scratch = buffers[1];
        init_frontier(kcore_problem);              // (the full list is never overwritten: one upload serves every k)
        int inside = num_nodes;                    // vertices of degree >= k after the last peel round
        for (int k = 1; k <= num_nodes; ++k) {
            for (;;) {
                const int peeled = filter_kernel<kcore_problem_t, deg_less_than_k_functor_t>(kcore_problem, everyone, scratch, k, context);
                if (peeled == 0) break;
                advance_forward_kernel<kcore_problem_t, update_deg_functor_t, /*idempotence=*/false, /*has_output=*/false>(
                    kcore_problem, scratch, everyone, k, context);
                inside = filter_kernel<kcore_problem_t, deg_atleast_k_functor_t>(kcore_problem, everyone, scratch, k, context);
            }
            if (inside == 0) {
                std::cout << "largest k-core: " << k - 1 << std::endl;
                kcore_problem->largest_k_core = k - 1;
                break;
            }
        }
    }
};

}  // namespace kcore
}  // namespace gunrock
