// coloring_functor.hxx -- the functors of the hash-extrema colouring (names and argument lists of
// gunrock/src/coloring/coloring_functor.hxx:10-70).
#pragma once
#include <climits>
#include "coloring/coloring_problem.hxx"
#include "intrinsics.hxx"

using namespace gunrock::util;

namespace gunrock {
namespace coloring {

typedef coloring_problem_t::data_slice_t coloring_slice_t;

// filter: a vertex whose hash is below every uncoloured neighbour's takes colour 2*it+1, one above every
// uncoloured neighbour's takes 2*it+2; it leaves the frontier once coloured
struct coloring_functor_t {
    GUNROCK_FN bool cond_filter(GUNROCK_VERTEX_ARGS(coloring_slice_t)) {
        const int h = data->d_hashs[idx];
        int color = 0;
        if (h < data->d_reduced_min[idx]) color = 2 * iteration + 1;
        else if (h > data->d_reduced_max[idx]) color = 2 * iteration + 2;
        if (color == 0) return true;
        data->d_colors[idx] = color;
        return false;
    }
};

// what a neighbour contributes to the extrema: its hash while it is uncoloured, else the reduction's neutral value
template <bool MAX>
struct reduce_hash_t {
    GUNROCK_FN bool cond_advance(GUNROCK_ARC_ARGS(coloring_slice_t)) { return true; }
    GUNROCK_FN bool apply_advance(GUNROCK_ARC_ARGS(coloring_slice_t)) { return true; }
    GUNROCK_FN int get_value_to_reduce(GUNROCK_VERTEX_ARGS(coloring_slice_t)) {
        if (data->d_colors[idx] == 0) return data->d_hashs[idx];
        return MAX ? INT_MIN : INT_MAX;   // (numeric_limits is host-only without --expt-relaxed-constexpr)
    }
};
typedef reduce_hash_t<true> reduce_max_t;
typedef reduce_hash_t<false> reduce_min_t;

}  // namespace coloring
}  // namespace gunrock
