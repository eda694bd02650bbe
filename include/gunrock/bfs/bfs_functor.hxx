// bfs_functor.hxx -- the BFS functor set (names and argument lists of
// gunrock/src/bfs/bfs_functor.hxx:7-53).
#pragma once
#include "bfs/bfs_problem.hxx"

namespace gunrock {
namespace bfs {

struct bfs_functor_t {
    typedef bfs_problem_t::data_slice_t slice_t;

    // filter: drop the -1 holes an un-compacted advance leaves -- and nothing else, which filter_kernel may rely on
    // (filter.hxx, detail::filter_drops_only_holes) when its input is known to have none
    static constexpr bool cond_filter_drops_only_holes = true;
    GUNROCK_FN bool cond_filter(GUNROCK_VERTEX_ARGS(slice_t)) { return idx != -1; }

    // uniquify: label a vertex the first time it is seen.  Valid for every source and for
    // vertex 0 (the reference's version only accepts idx > 0 and labels > 0);
    // exactness comes from the visited-bit test-and-set that precedes this call.
    GUNROCK_FN bool cond_uniq(GUNROCK_VERTEX_ARGS(slice_t)) {
        if (idx < 0) return false;
        const int seen = data->d_labels[idx];
        if (seen >= 0 && seen <= iteration) return false;
        data->d_labels[idx] = iteration + 1;
        return true;
    }

    // advance: visit dst if nobody has yet; the atomic decides the winner, who writes the depth.  (The reference tests
    // `d_labels[dst] == -1` and claims with atomicCAS on the label, bfs_functor.hxx:26-33; here the visited BIT of the
    // problem's slice is tested and claimed and the winner stores the label -- same labels, 32x smaller probe target.)
    // cond_advance is a pure read and apply_advance can only succeed where it holds, so the engine may probe with the
    // one and commit with the other (advance.hxx, detail::cond_guards_apply): ~4 M atomics per scale-22 traversal instead
    // of one per arc (134 M).
    static constexpr bool cond_advance_guards_apply = true;
    GUNROCK_FN bool cond_advance(GUNROCK_ARC_ARGS(slice_t)) {
        return !((data->d_visited[(unsigned)dst >> 5] >> (dst & 31)) & 1u);
    }
    GUNROCK_FN bool apply_advance(GUNROCK_ARC_ARGS(slice_t)) {
        const unsigned bit = 1u << (dst & 31);
        if (atomicOr(data->d_visited + ((unsigned)dst >> 5), bit) & bit) return false;
        data->d_labels[dst] = iteration + 1;
        return true;
    }

    // push -> pull hand-over
    GUNROCK_FN bool cond_sparse_to_dense(GUNROCK_VERTEX_ARGS(slice_t)) {
        return data->d_labels[idx] == iteration;
    }
    GUNROCK_FN bool cond_gen_unvisited(GUNROCK_VERTEX_ARGS(slice_t)) {
        return data->d_labels[idx] == -1;
    }

    // neighborhood-reduce hooks (unused by BFS itself)
    GUNROCK_FN int get_value_to_reduce(GUNROCK_VERTEX_ARGS(slice_t)) { return iteration; }
    GUNROCK_FN void write_reduced_value(int item, int val, slice_t *data, int iteration) {}
};

}  // namespace bfs
}  // namespace gunrock
