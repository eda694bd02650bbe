// neighborhood.hxx -- gunrock::oprtr::neighborhood::neighborhood_kernel with the
// reference signature (gunrock/src/neighborhood.hxx:12-19) on the engine's
// warp-segmented reduce (include/b200/segreduce.cuh) instead of mgpu lbs_segreduce.
//
// reduced[slot] = reduce_op over the neighbours u of input[slot] of
// Functor::get_value_to_reduce(u); `identity` for an empty neighbourhood.  The optional
// trailing template flag write_back = true additionally calls
// Functor::write_reduced_value(vertex, value) per slot -- the scatter the reference left
// commented out (neighborhood.hxx:60-67).
#pragma once
#include "b200/operators.cuh"
#include "frontier.hxx"
#include "intrinsics.hxx"

namespace gunrock {
namespace oprtr {
namespace neighborhood {

namespace detail {
template <typename Value, typename reduce_op> struct rop_for;
template <> struct rop_for<float, mgpu::plus_t<float>> { using type = b200::PlusF32; };
template <> struct rop_for<float, mgpu::minimum_t<float>> { using type = b200::MinF32; };
template <> struct rop_for<float, mgpu::maximum_t<float>> { using type = b200::MaxF32; };
template <> struct rop_for<int, mgpu::plus_t<int>> { using type = b200::PlusI32; };
template <> struct rop_for<int, mgpu::minimum_t<int>> { using type = b200::MinI32; };
template <> struct rop_for<int, mgpu::maximum_t<int>> { using type = b200::MaxI32; };

// ADVANCE_HERE: evaluate cond_advance / apply_advance in this pass.  The reference evaluates both advance halves ONCE
// per arc before taking the value (neighborhood.hxx:50-54); with has_output the emit pass (EmitOp) already did, and a
// functor with side effects (atomics, counters) must not see every arc twice.
template <typename Problem, typename Functor, typename Value, bool ADVANCE_HERE>
struct ValueOfNeighbor {
    typename Problem::data_slice_t *data;
    int iteration;
    __device__ __forceinline__ Value operator()(int src, int nbr, uint32_t eid) const {
        if (ADVANCE_HERE) {
            (void)Functor::cond_advance(src, nbr, (int)eid, 0, (int)eid, data, iteration);
            (void)Functor::apply_advance(src, nbr, (int)eid, 0, (int)eid, data, iteration);
        }
        return Functor::get_value_to_reduce(nbr, data, iteration);
    }
};

template <typename Problem, typename Functor, bool has_output>
struct EmitOp {   // has_output = true: out[idx] = (cond && apply) ? nbr : -1 (neighborhood.hxx:52-54)
    typename Problem::data_slice_t *data;
    int iteration;
    __device__ __forceinline__ bool probe(int, int, uint32_t) const { return true; }
    __device__ __forceinline__ bool commit(int src, int dst, uint32_t eid, uint32_t rank, uint32_t out_idx) const {
        const bool c = Functor::cond_advance(src, dst, (int)eid, (int)rank, (int)out_idx, data, iteration);
        const bool a = Functor::apply_advance(src, dst, (int)eid, (int)rank, (int)out_idx, data, iteration);
        return c && a;
    }
};

// write_reduced_value scatter, compiled only when asked for (not every Functor defines it)
template <typename Problem, typename Functor, typename Value>
struct WriteBackFn {
    const int *in;
    const Value *reduced;
    typename Problem::data_slice_t *data;
    int iteration;
    __device__ void operator()(int slot) const { Functor::write_reduced_value(in[slot], reduced[slot], data, iteration); }
};
template <typename Problem, typename Functor, typename Value>
void write_back(std::true_type, const int *in, const Value *reduced, typename Problem::data_slice_t *data, int iteration,
                size_t len, standard_context_t &context) {
    transform(WriteBackFn<Problem, Functor, Value>{in, reduced, data, iteration}, len, context);
}
template <typename Problem, typename Functor, typename Value>
void write_back(std::false_type, const int *, const Value *, typename Problem::data_slice_t *, int, size_t,
                standard_context_t &) {}

template <typename Problem, typename Functor>
void emit_raw(std::true_type, b200_workspace *ws, const b200::LbsArgs &a, typename Problem::data_slice_t *data,
              int iteration, std::shared_ptr<frontier_t<int>> &output) {
    EmitOp<Problem, Functor, true> op{data, iteration};
    mgpu::throw_on_error(b200::launch_lbs_advance<b200::OUT_RAW, false>(ws, a, op, output->data()->data(), output->capacity()));
}
template <typename Problem, typename Functor>
void emit_raw(std::false_type, b200_workspace *, const b200::LbsArgs &, typename Problem::data_slice_t *, int,
              std::shared_ptr<frontier_t<int>> &) {}
}  // namespace detail

// (push defaults to false = pull over the CSC: the reference's coloring / lspar call sites pass five template
// arguments, coloring_enactor.hxx:60-76, one short of its own declaration)
template <typename Problem, typename Functor, typename Value, typename reduce_op, bool has_output, bool push = false,
          bool write_back = false>
int neighborhood_kernel(std::shared_ptr<Problem> problem, std::shared_ptr<frontier_t<int>> &input,
                        std::shared_ptr<frontier_t<int>> &output, Value *reduced, Value identity, int iteration,
                        standard_context_t &context) {
    using ROp = typename detail::rop_for<Value, reduce_op>::type;
    const size_t len = input->size();
    if (!len) return 0;
    graph_device_t &g = *problem->gslice;
    if (b200_ctx_reserve(context.engine(), (int64_t)len) != B200_OK) throw cuda_exception_t(cudaErrorMemoryAllocation);
    b200_workspace *ws = context.workspace();
    const int *in = input->data()->data();
    const uint32_t *offsets = reinterpret_cast<const uint32_t *>(push ? g.d_row_offsets.data() : g.d_col_offsets.data());
    const int *indices = push ? g.d_col_indices.data() : g.d_row_indices.data();
    typename Problem::data_slice_t *data = problem->d_data_slice.data();

    mgpu::throw_on_error(b200::reset_counters(ws));
    // degree scan; presets reduced[slot] to identity (empty) or the operator's neutral element
    b200::NeighborhoodDegree<Value> deg{in, offsets, reduced, identity, ROp::neutral(), 0};
    mgpu::throw_on_error(b200::launch_scan(ws, deg, (uint32_t)len, ws->d_scanned, ws->d_counters + B200_CNT_TOTAL));
    const b200::LbsArgs a = b200::make_lbs_args(ws, in, (uint32_t)len, offsets, indices);
    if (has_output) {   // the emit pass writes one entry per arc: make room first (resize prints the reference's overflow text)
        mgpu::throw_on_error(b200::read_counters(ws));
        output->resize((size_t)ws->h_counters[B200_CNT_TOTAL]);
        output->set_hole_free(false);   // raw layout: -1 where the functor rejected the arc
    }
    detail::emit_raw<Problem, Functor>(std::integral_constant<bool, has_output>(), ws, a, data, iteration, output);
    detail::ValueOfNeighbor<Problem, Functor, Value, !has_output> vf{data, iteration};
    mgpu::throw_on_error((b200::launch_lbs_segreduce<Value, ROp>(ws, a, vf, reduced, 0)));
    detail::write_back<Problem, Functor, Value>(std::integral_constant<bool, write_back>(), in, reduced, data, iteration, len, context);
    mgpu::throw_on_error(b200::read_counters(ws));
    const int non_zeros = (int)ws->h_counters[B200_CNT_TOTAL];
    if (has_output && non_zeros) output->resize((size_t)non_zeros);
    return non_zeros;
}

}  // namespace neighborhood
}  // namespace oprtr
}  // namespace gunrock
