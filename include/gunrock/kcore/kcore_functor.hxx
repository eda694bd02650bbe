// kcore_functor.hxx -- the three functors of the k-core peel (names and argument lists of
// gunrock/src/kcore/kcore_functor.hxx:10-37).  `iteration` carries k.
#pragma once
#include "intrinsics.hxx"
#include "kcore/kcore_problem.hxx"

using namespace gunrock::util;

namespace gunrock {
namespace kcore {

typedef kcore_problem_t::data_slice_t kcore_slice_t;

// filter: the vertices that fall out of the k-core now; records their core number and retires them
struct deg_less_than_k_functor_t {
    GUNROCK_FN bool cond_filter(GUNROCK_VERTEX_ARGS(kcore_slice_t)) {
        const int k = iteration, d = data->d_degrees[idx];
        if (d <= 0 || d >= k) return false;
        data->d_num_cores[idx] = k - 1;
        data->d_degrees[idx] = 0;
        return true;
    }
};

// filter: the vertices still inside the k-core
struct deg_atleast_k_functor_t {
    GUNROCK_FN bool cond_filter(GUNROCK_VERTEX_ARGS(kcore_slice_t)) { return data->d_degrees[idx] >= iteration; }
};

// advance over the retired vertices (has_output = false): every neighbour loses one degree
struct update_deg_functor_t {
    GUNROCK_FN bool cond_advance(GUNROCK_ARC_ARGS(kcore_slice_t)) { return true; }
    GUNROCK_FN bool apply_advance(GUNROCK_ARC_ARGS(kcore_slice_t)) {
        return atomicSub(data->d_degrees + dst, 1) > 1;   // (the reference re-reads the degree after its atomicAdd(-1))
    }
};

}  // namespace kcore
}  // namespace gunrock
