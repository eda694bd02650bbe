// bfs_enactor.hxx -- BFS enactor: push levels (advance + filter) then, once
// unvisited < |frontier| * threshold, pull levels until nothing changes.  Control flow,
// printed lines and the meaning of `threshold` follow gunrock/src/bfs/bfs_enactor.hxx:41-117.
// enact_builtin() runs the same traversal inside the engine (b200_bfs_run): bitmap
// frontier, early-exit pull, no per-operator host round trip.
#pragma once
#include "advance.hxx"
#include "bfs_functor.hxx"
#include "bfs_problem.hxx"
#include "enactor.hxx"
#include "filter.hxx"
#include "frontier.hxx"
#include "graph.hxx"
#include "test_utils.hxx"

using namespace mgpu;
using namespace gunrock::oprtr::advance;
using namespace gunrock::oprtr::filter;

namespace gunrock {
namespace bfs {

struct bfs_enactor_t : enactor_t {
    using problem_ptr = std::shared_ptr<bfs_problem_t>;
    using P = bfs_problem_t;
    using F = bfs_functor_t;

    bfs_enactor_t(standard_context_t &context, int num_nodes, int num_edges) : enactor_t(context, num_nodes, num_edges) {}

    // the frontier of level 0 is the source alone
    void init_frontier(problem_ptr bfs_problem) { buffers[0]->load(std::vector<int>(1, bfs_problem->src)); }

    void enact_pushpull(problem_ptr bfs_problem, float threshold, standard_context_t &context) {
        init_frontier(bfs_problem);
        const int num_nodes = bfs_problem->gslice->num_nodes;
        int remaining = num_nodes - 1;   // vertices without a label
        int frontier_length = 1, cur = 0, iteration = 0;

        // ---- push: expand the frontier, keep the newly labelled vertices
        for (;; ++iteration) {
            frontier_length = advance_forward_kernel<P, F, false, true>(bfs_problem, buffers[cur], buffers[cur ^ 1], iteration, context);
            cur ^= 1;
            if (!frontier_length) break;
            frontier_length = filter_kernel<P, F>(bfs_problem, buffers[cur], buffers[cur ^ 1], iteration, context);
            remaining -= frontier_length;
            if ((float)remaining < (float)frontier_length * threshold) break;   // hand over to pull
            if (!frontier_length) break;
            cur ^= 1;
        }
        std::cout << "pushed iterations: " << iteration << std::endl;
        if (!frontier_length) return;                                            // traversal finished while pushing

        // ---- pull: unvisited vertices look for a parent in the current frontier
        ++iteration;
        frontier_length = gen_unvisited_kernel<P, F>(bfs_problem, unvisited[cur ^ 1], unvisited[cur], 0, context);
        mem_t<int> zeros = mgpu::fill<int>(0, num_nodes, context);
        buffers[cur]->load(zeros);
        sparse_to_dense_kernel<P, F>(bfs_problem, buffers[cur ^ 1], buffers[cur], iteration, context);
        for (;; ++iteration) {
            buffers[cur ^ 1]->load(zeros);
            advance_backward_kernel<P, F>(bfs_problem, unvisited[cur], buffers[cur], buffers[cur ^ 1], iteration, context);
            const int still_unvisited = filter_kernel<P, F>(bfs_problem, unvisited[cur], unvisited[cur ^ 1], iteration, context);
            if (!still_unvisited || still_unvisited == frontier_length) break;
            frontier_length = still_unvisited;
            cur ^= 1;
        }
        std::cout << "total iterations: " << iteration << std::endl;
    }

    // Whole traversal inside the engine; labels land in bfs_problem->d_labels.
    // mode: B200_BFS_PUSH / B200_BFS_REF_ALPHA (threshold = alpha) / B200_BFS_BEAMER.
    int enact_builtin(problem_ptr bfs_problem, int mode, float alpha, float beta, standard_context_t &context,
                      b200_stats *stats = nullptr) {
        const b200_graph g = bfs_problem->gslice->view();
        return b200_bfs_run(context.engine(), &g, bfs_problem->src, mode, alpha, beta, bfs_problem->d_labels.data(), stats);
    }
};

}  // namespace bfs
}  // namespace gunrock
