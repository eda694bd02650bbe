// sssp_enactor.hxx -- frontier-driven Bellman-Ford: relax the frontier's out-arcs, keep
// each improved vertex once, repeat until no distance changes
// (gunrock/src/sssp/sssp_enactor.hxx:40-72).  enact_builtin() runs it inside the engine
// (b200_sssp_run), which also produces deterministic predecessors.
#pragma once
#include "advance.hxx"
#include "enactor.hxx"
#include "filter.hxx"
#include "frontier.hxx"
#include "graph.hxx"
#include "sssp_functor.hxx"
#include "sssp_problem.hxx"
#include "test_utils.hxx"

using namespace mgpu;
using namespace gunrock::oprtr::advance;
using namespace gunrock::oprtr::filter;

namespace gunrock {
namespace sssp {

struct sssp_enactor_t : enactor_t {
    sssp_enactor_t(standard_context_t &context, int num_nodes, int num_edges, float queue_sizing)
        : enactor_t(context, num_nodes, num_edges, queue_sizing) {}
    sssp_enactor_t(const sssp_enactor_t &) = delete;
    sssp_enactor_t &operator=(const sssp_enactor_t &) = delete;

    void init_frontier(std::shared_ptr<sssp_problem_t> sssp_problem) {
        buffers[0]->load(std::vector<int>(1, sssp_problem->src));
    }

    void enact(std::shared_ptr<sssp_problem_t> sssp_problem, standard_context_t &context) {
        init_frontier(sssp_problem);
        int cur = 0;
        for (int iteration = 0;; ++iteration) {
            advance_forward_kernel<sssp_problem_t, sssp_functor_t, false, true>(sssp_problem, buffers[cur], buffers[cur ^ 1],
                                                                                iteration, context);
            cur ^= 1;
            const int improved = filter_kernel<sssp_problem_t, sssp_functor_t>(sssp_problem, buffers[cur], buffers[cur ^ 1],
                                                                               iteration, context);
            if (!improved) break;
            cur ^= 1;
        }
    }

    int enact_builtin(std::shared_ptr<sssp_problem_t> sssp_problem, standard_context_t &context, b200_stats *stats = nullptr) {
        const b200_graph g = sssp_problem->gslice->view();
        return b200_sssp_run(context.engine(), &g, sssp_problem->src, sssp_problem->d_labels.data(),
                             sssp_problem->d_preds.data(), stats);
    }
};

}  // namespace sssp
}  // namespace gunrock
