// sssp_functor.hxx -- the SSSP functor set (names and argument lists of
// gunrock/src/sssp/sssp_functor.hxx:10-36): relax with a float atomicMin, stamp-based
// frontier dedupe.
#pragma once
#include "intrinsics.hxx"
#include "sssp/sssp_problem.hxx"

using namespace gunrock::util;

namespace gunrock {
namespace sssp {

struct sssp_functor_t {
    typedef sssp_problem_t::data_slice_t slice_t;

    // keep a vertex once per iteration: the stamp test-and-set is one atomic exchange here
    // (the reference reads then writes d_visited non-atomically and may keep duplicates)
    static __device__ __forceinline__ bool cond_filter(int idx, slice_t *data, int iteration) {
        if (idx == -1) return false;
        return atomicExch(data->d_visited + idx, iteration) != iteration;
    }

    // relax src -> dst; true when this thread lowered dst's distance
    static __device__ __forceinline__ bool cond_advance(int src, int dst, int edge_id, int rank, int output_idx,
                                                        slice_t *data, int iteration) {
        const float through_src = data->d_labels[src] + data->d_weights[edge_id];
        return through_src < atomicMin(data->d_labels + dst, through_src);
    }

    // record the tail as predecessor (last writer wins, as in the reference)
    static __device__ __forceinline__ bool apply_advance(int src, int dst, int edge_id, int rank, int output_idx,
                                                         slice_t *data, int iteration) {
        data->d_preds[dst] = src;
        return true;
    }
};

}  // namespace sssp
}  // namespace gunrock
