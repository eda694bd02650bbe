// coloring_problem.hxx -- graph colouring problem data with the reference's member names
// (gunrock/src/coloring/coloring_problem.hxx:8-62): a random hash per vertex (redrawn every iteration),
// the neighbourhood extrema of the hashes, and the colours (0 = not coloured yet).
#pragma once
#include <random>
#include "problem.hxx"

namespace gunrock {
namespace coloring {

struct coloring_problem_t : problem_t {
    struct data_slice_t {
        int *d_reduced_max, *d_reduced_min, *d_hashs, *d_colors;
    };

    mem_t<int> d_reduced_max, d_reduced_min, d_hashs, d_colors;
    int prime = 0;       // hashes are uniform in [0, prime]
    int max_iter = 0;
    mem_t<data_slice_t> d_data_slice;

    coloring_problem_t() = default;

    coloring_problem_t(std::shared_ptr<graph_device_t> graph, int prime_, int iterations, standard_context_t &context)
        : problem_t(graph), prime(prime_), max_iter(iterations) {
        const int n = graph->num_nodes;
        d_reduced_max = mgpu::fill(0, n, context);
        d_reduced_min = mgpu::fill(0, n, context);
        d_colors = mgpu::fill(0, n, context);
        reset_hashs(context);
    }

    // New hashes for the vertices (all of them, coloured or not, as in the reference) from one process-wide
    // std::mt19937 in its default state -- what mgpu::fill_random(0, prime, n, false, ctx) draws (memory.hxx:112-129),
    // so a run sees the reference's hash sequence.  The slice is re-published because d_hashs moves.
    void reset_hashs(standard_context_t &context) {
        static std::mt19937 engine;
        std::uniform_int_distribution<int> draw(0, prime);
        std::vector<int> h(gslice->num_nodes);
        for (int &x : h) x = draw(engine);
        d_hashs = to_mem(h, context);
        d_data_slice = publish_slice(data_slice_t{d_reduced_max.data(), d_reduced_min.data(), d_hashs.data(), d_colors.data()}, context);
    }

    std::vector<int> colors;
    void extract() { colors = from_mem(d_colors); }   // (the reference's is a stub; its test prints d_colors instead)
};

}  // namespace coloring
}  // namespace gunrock
