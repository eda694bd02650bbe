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
    GUNROCK_FN bool cond_filter(GUNROCK_VERTEX_ARGS(slice_t)) {
        if (idx == -1) return false;
        return atomicExch(data->d_visited + idx, iteration) != iteration;
    }

    // relax src -> dst; true when this thread lowered dst's distance
    GUNROCK_FN bool cond_advance(GUNROCK_ARC_ARGS(slice_t)) {
        const float through_src = data->d_labels[src] + data->d_weights[edge_id];
        return through_src < atomicMin(data->d_labels + dst, through_src);
    }

    // record the tail as predecessor (last writer wins, as in the reference)
    GUNROCK_FN bool apply_advance(GUNROCK_ARC_ARGS(slice_t)) {
        data->d_preds[dst] = src;
        return true;
    }
};

}  // namespace sssp
}  // namespace gunrock
