// problem.hxx -- problem_t: base of every primitive's problem; holds the graph slice
// (gunrock/src/problem.hxx:7-31).
#pragma once
#include "graph.hxx"

namespace gunrock {

struct problem_t {
    std::shared_ptr<graph_device_t> gslice;

    problem_t() : gslice(std::make_shared<graph_device_t>()) {}
    explicit problem_t(std::shared_ptr<graph_device_t> rhs) : gslice(rhs) {}
    problem_t(const problem_t &) = delete;
    problem_t &operator=(const problem_t &) = delete;

    // out-degree of every vertex as float (problem.hxx:23-30)
    void GetDegrees(mem_t<float> &_degrees, standard_context_t &context) {
        float *deg = _degrees.data();
        const int *off = gslice->d_row_offsets.data();
        transform([=] __device__(int v) { deg[v] = (float)(off[v + 1] - off[v]); }, gslice->num_nodes, context);
    }
};

}  // namespace gunrock
