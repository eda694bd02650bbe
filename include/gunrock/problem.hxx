// problem.hxx -- problem_t: base of every primitive's problem; holds the graph slice
// (gunrock/src/problem.hxx:7-31).
#pragma once
#include "graph.hxx"

// Parameter lists of the functor contract (advance.hxx:47-59, filter.hxx:21-24,
// neighborhood.hxx:47-58 call them in exactly this form), so that a functor states only what it computes:
//   GUNROCK_FN bool cond_filter(GUNROCK_VERTEX_ARGS(slice_t)) { return idx != -1; }
#define GUNROCK_FN static __device__ __forceinline__
#define GUNROCK_VERTEX_ARGS(slice_type) int idx, slice_type *data, int iteration
#define GUNROCK_ARC_ARGS(slice_type) \
    int src, int dst, int edge_id, int rank, int output_idx, slice_type *data, int iteration

namespace gunrock {

struct problem_t {
    std::shared_ptr<graph_device_t> gslice;

    problem_t() : gslice(std::make_shared<graph_device_t>()) {}
    explicit problem_t(std::shared_ptr<graph_device_t> rhs) : gslice(rhs) {}
    problem_t(const problem_t &) = delete;
    problem_t &operator=(const problem_t &) = delete;

    // Uploads the POD a primitive's functors receive (Problem::data_slice_t) and returns the device
    // copy that becomes Problem::d_data_slice.  (The reference spells this out in every problem as a host vector of
    // one slice plus an init() member.)
    template <class Slice>
    static mem_t<Slice> publish_slice(const Slice &slice, standard_context_t &context) {
        std::vector<Slice> staged(1, slice);
        return to_mem(staged, context);
    }

    // out-degree of every vertex as float (problem.hxx:23-30)
    void GetDegrees(mem_t<float> &_degrees, standard_context_t &context) {
        float *deg = _degrees.data();
        const int *off = gslice->d_row_offsets.data();
        transform([=] __device__(int v) { deg[v] = (float)(off[v + 1] - off[v]); }, gslice->num_nodes, context);
    }
};

}  // namespace gunrock
