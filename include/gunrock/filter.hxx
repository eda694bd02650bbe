// filter.hxx -- gunrock::oprtr::filter with the reference signatures
// (gunrock/src/filter.hxx:11-16, 95-101) on the engine's single-pass stable compaction
// (include/b200/tile_scan.cuh) instead of mgpu transform_compact's upsweep / scan /
// pinned read-back / downsweep.
#pragma once
#include <type_traits>
#include "b200/operators.cuh"
#include "frontier.hxx"

namespace gunrock {
namespace oprtr {
namespace filter {

namespace detail {
template <typename Problem, typename Functor>
struct CondFilterPred {
    const int *items;
    typename Problem::data_slice_t *data;
    int iteration;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = items[idx];
        return Functor::cond_filter(item, data, iteration);   // runs once per item; may mutate problem data
    }
};

// Exact duplicate removal: atomic test-and-set of the item's bit in the visited mask
// (one bit per vertex, addressed as item>>3 / item&7 like the reference's mask), then
// the problem's cond_uniq.  Replaces the bitmask / warp-hash / history-hash heuristics of
// UniquifyFunctor (filter.hxx:33-91), which may let duplicates through, skip item 0 and
// mis-mask (1 << item & 7); here every vertex passes exactly once, for any source.
template <typename Problem, typename ProblemFunctor>
struct UniquifyPred {
    const int *items;
    unsigned char *mask;
    typename Problem::data_slice_t *data;
    int iteration;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = items[idx];
        if (item < 0) return false;
        const size_t byte = (size_t)item >> 3;
        unsigned int *word = reinterpret_cast<unsigned int *>(mask + (byte & ~(size_t)3));
        const unsigned int bit = (1u << (item & 7)) << (8 * (byte & 3));
        if ((*word & bit) || (atomicOr(word, bit) & bit)) return false;
        return ProblemFunctor::cond_uniq(item, data, iteration);
    }
};

template <typename Pred>
int run_compact(Pred pred, size_t len, std::shared_ptr<frontier_t<int>> &output, standard_context_t &context) {
    output->set_hole_free(true);
    if (!len) {
        output->resize(0);
        return 0;
    }
    if (b200_ctx_reserve(context.engine(), (int64_t)len) != B200_OK) throw cuda_exception_t(cudaErrorMemoryAllocation);
    b200_workspace *ws = context.workspace();
    mgpu::throw_on_error(b200::reset_counters(ws));
    mgpu::throw_on_error(b200::launch_compact(ws, pred, (uint32_t)len, output->data()->data(), output->capacity(),
                                              ws->d_counters + B200_CNT_OUT, ws->d_counters + B200_CNT_OVERFLOW));
    mgpu::throw_on_error(b200::read_counters(ws));
    output->resize((size_t)ws->h_counters[B200_CNT_OUT]);
    return (int)output->size();
}
// Opt-in functor trait: `static constexpr bool cond_filter_drops_only_holes = true;` states that cond_filter is
// exactly `item != -1` with no side effect (bfs_functor.hxx:9-11: its only job is to drop the holes advance.hxx:60
// leaves).  On a frontier the operators know to be hole-free (frontier_t::hole_free) such a filter keeps everything:
// the list is copied on the stream and its length is already known on the host -- no kernel pass, no read-back.
template <typename F, typename = void>
struct filter_drops_only_holes : std::false_type {};
template <typename F>
struct filter_drops_only_holes<F, std::void_t<decltype(F::cond_filter_drops_only_holes)>>
    : std::integral_constant<bool, F::cond_filter_drops_only_holes> {};
}  // namespace detail

template <typename Problem, typename Functor>
int filter_kernel(std::shared_ptr<Problem> problem, std::shared_ptr<frontier_t<int>> &input,
                  std::shared_ptr<frontier_t<int>> &output, int iteration, standard_context_t &context) {
    if (detail::filter_drops_only_holes<Functor>::value && input->hole_free()) {
        const size_t len = input->size();
        output->resize(len);   // (prints the reference's overflow message and exits if it does not fit)
        if (len)
            mgpu::throw_on_error(cudaMemcpyAsync(output->data()->data(), input->data()->data(), sizeof(int) * len,
                                                 cudaMemcpyDeviceToDevice, context.stream()));
        output->set_hole_free(true);
        return (int)len;
    }
    detail::CondFilterPred<Problem, Functor> pred{input->data()->data(), problem->d_data_slice.data(), iteration};
    return detail::run_compact(pred, input->size(), output, context);
}

// d_visited_mask: at least ceil(num_nodes / 8) bytes, rounded up to a multiple of 4, zeroed by
// the caller before the first level.
template <typename Problem, typename ProblemFunctor>
void uniquify_kernel(std::shared_ptr<Problem> problem, unsigned char *d_visited_mask,
                     std::shared_ptr<frontier_t<int>> &input, std::shared_ptr<frontier_t<int>> &output, int iteration,
                     standard_context_t &context) {
    detail::UniquifyPred<Problem, ProblemFunctor> pred{input->data()->data(), d_visited_mask,
                                                       problem->d_data_slice.data(), iteration};
    detail::run_compact(pred, input->size(), output, context);
}

}  // namespace filter
}  // namespace oprtr
}  // namespace gunrock
