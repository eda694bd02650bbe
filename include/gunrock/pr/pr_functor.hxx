// pr_functor.hxx -- the PageRank-style functor set (names and argument lists of
// gunrock/src/pr/pr_functor.hxx:10-30).
#pragma once
#include "intrinsics.hxx"
#include "pr/pr_problem.hxx"

using namespace gunrock::util;

namespace gunrock {
namespace pr {

struct pr_functor_t {
    typedef pr_problem_t::data_slice_t slice_t;

    // rank update + convergence test: keep the vertex while its rank still moves by > 0.1 %
    GUNROCK_FN bool cond_filter(GUNROCK_VERTEX_ARGS(slice_t)) {
        const float before = data->d_current_ranks[idx];
        const float deg = data->d_degrees[idx];
        float after = 0.15f;
        if (deg > 0) after = 0.15f + 0.85f * data->d_reduced_ranks[idx] / deg;
        if (!isfinite(after)) after = 0;
        data->d_current_ranks[idx] = after;
        return fabs(after - before) > (0.001f * before);
    }

    GUNROCK_FN bool cond_advance(GUNROCK_ARC_ARGS(slice_t)) { return true; }
    GUNROCK_FN bool apply_advance(GUNROCK_ARC_ARGS(slice_t)) { return true; }

    // a neighbour contributes its current rank (non-finite ranks count as 0)
    GUNROCK_FN float get_value_to_reduce(GUNROCK_VERTEX_ARGS(slice_t)) {
        const float r = data->d_current_ranks[idx];
        return isfinite(r) ? r : 0.0f;
    }
    // scatter hook for neighborhood_kernel<..., write_back = true>
    GUNROCK_FN void write_reduced_value(int item, float val, slice_t *data, int iteration) {
        data->d_reduced_ranks[item] = val;
    }
};

}  // namespace pr
}  // namespace gunrock
