// enactor.hxx -- enactor_t: ping-pong frontier buffers plus the iota / unvisited index
// arrays every enactor starts from (gunrock/src/enactor.hxx:11-46).
#pragma once
#include "frontier.hxx"
#include "graph.hxx"
#include "problem.hxx"

using namespace mgpu;

namespace gunrock {

struct enactor_t {
    using frontier_ptr = std::shared_ptr<frontier_t<int>>;
    std::vector<frontier_ptr> buffers;      // ping-pong frontiers
    frontier_ptr indices, filtered_indices; // iota(n) and its filtered twin ...
    std::vector<frontier_ptr> unvisited;    // ... as the ping-pong pair the pull phase walks

    enactor_t(standard_context_t &context, int num_nodes, int num_edges, float queue_sizing = 1.0f) {
        init(context, num_nodes, num_edges, queue_sizing);
    }
    enactor_t(const enactor_t &) = delete;
    enactor_t &operator=(const enactor_t &) = delete;

    void init(standard_context_t &context, int num_nodes, int num_edges, float queue_sizing) {
        // capacity as in the reference (num_edges * queue_sizing) so RAW-layout advances fit;
        // never below num_nodes because the pull phase stores per-vertex flags in these buffers
        const size_t cap = std::max((size_t)((double)num_edges * queue_sizing), (size_t)num_nodes);
        auto make = [&](size_t capacity) { return std::make_shared<frontier_t<int>>(context, capacity); };
        buffers = {make(cap), make(cap)};
        indices = make(num_nodes);
        filtered_indices = make(num_nodes);
        mem_t<int> iota = fill_function<int>([] __device__(int i) { return i; }, num_nodes, context);
        indices->load(iota);
        filtered_indices->load(iota);
        unvisited = {indices, filtered_indices};
        b200_ctx_reserve(context.engine(), (int64_t)cap);
    }
};

}  // namespace gunrock
