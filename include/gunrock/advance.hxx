// advance.hxx -- the advance operators of gunrock::oprtr::advance with the reference
// signatures (gunrock/src/advance.hxx:20-25, 69-74, 86-91, 108-114), implemented on the
// B200 engine kernels (include/b200/*.cuh) instead of mgpu transform_scan + transform_lbs
// + transform_compact.  User functors are instantiated into those kernels here, in the
// caller's translation unit.
//
// Output layout of advance_forward_kernel:
//   default                      compacted: only accepted neighbours, no -1 holes (advance
//                                + the `!= -1` filter fused); return value = items written.
//   -DB200_ADVANCE_RAW_OUTPUT    reference layout out[idx] = nbr | -1 for every arc,
//                                return value = m_F (advance.hxx:60,66).
// Either way a following filter_kernel yields the same frontier, so enactors written
// against the reference work unchanged.
#pragma once
#include <type_traits>
#include "b200/operators.cuh"
#include "frontier.hxx"
#include "intrinsics.hxx"

namespace gunrock {
namespace oprtr {
namespace advance {

namespace detail {
inline void check(cudaError_t e) { mgpu::throw_on_error(e); }

// Adapts a reference-style Functor (static cond_advance/apply_advance taking the
// problem's data_slice_t*) to the engine's probe/commit protocol.  Both halves are
// always evaluated, exactly like advance.hxx:57-58.
template <typename Problem, typename Functor, bool idempotence>
struct RefFunctorOp {
    typename Problem::data_slice_t *data;
    int iteration;
    __device__ __forceinline__ bool probe(int, int, uint32_t) const { return true; }
    __device__ __forceinline__ bool commit(int src, int dst, uint32_t eid, uint32_t rank, uint32_t out_idx) const {
        const bool cond = Functor::cond_advance(src, dst, (int)eid, (int)rank, (int)out_idx, data, iteration);
        const bool applied = Functor::apply_advance(src, dst, (int)eid, (int)rank, (int)out_idx, data, iteration);
        return idempotence ? true : (cond && applied);
    }
};

// The same functor contract on the quad kernel's two-stage protocol (include/b200/quad_advance.cuh): a user functor has
// no separable "cheap probe" half, so every arc inside its row becomes a stage-2 candidate (probe_eval = true) and
// cond_advance / apply_advance run there, densely, one candidate per lane -- both always evaluated, exactly like
// advance.hxx:57-58.  What the adapter gains over the arc-wise kernel is the quad walk: one segment search and one
// 128-bit index load per four arcs, warp-private windows, no CTA barrier.
// rank = the arc's position in its row; output_idx: the reference's position in the un-compacted output does not exist
// here (the output is compacted) -- the arc id is passed, which is also unique per arc.
template <typename Problem, typename Functor, bool idempotence>
struct RefFunctorQ {
    typename Problem::data_slice_t *data;
    int iteration;
    const int *offsets;
    static constexpr bool WEIGHTED = false;
    using SrcVal = b200::NoSrc;
    using Evidence = b200::NoSrc;
    using Token = bool;
    using Cand = b200::CandWords<3>;   // dst, src, arc id
    __device__ __forceinline__ SrcVal load_src(int) const { return SrcVal(); }
    __device__ __forceinline__ Evidence probe_load(bool, SrcVal, int, int, uint32_t) const { return Evidence(); }
    __device__ __forceinline__ bool probe_eval(Evidence, SrcVal, int, float) const { return true; }
    __device__ __forceinline__ Cand make_cand(SrcVal, int src, int dst, uint32_t eid, float) const {
        return Cand{{(uint32_t)dst, (uint32_t)src, eid}};
    }
    __device__ __forceinline__ Token claim(const Cand &c) const {
        const int src = (int)c.w[1], dst = (int)c.w[0], eid = (int)c.w[2];
        const int rank = eid - offsets[src];   // (dead code unless the functor reads it)
        const bool cond = Functor::cond_advance(src, dst, eid, rank, eid, data, iteration);
        const bool applied = Functor::apply_advance(src, dst, eid, rank, eid, data, iteration);
        return idempotence ? true : (cond && applied);
    }
    __device__ __forceinline__ int finish(Token accepted, const Cand &c) const { return accepted ? (int)c.w[0] : -1; }
};

// Opt-in functor trait: `static constexpr bool cond_advance_guards_apply = true;` states that cond_advance has no side
// effect and that apply_advance cannot succeed (and changes nothing) for an arc whose cond_advance is false -- the BFS
// pair "label still -1?" / atomicCAS(label, -1, depth) is the model case (bfs_functor.hxx:26-35).  For such a functor
// `cond && apply` (advance.hxx:57-60) may short-circuit: cond_advance becomes the quad kernel's stage-1 probe (one
// cached load per arc, issued a pipeline stage ahead) and only the survivors reach apply_advance in stage 2, instead of
// one atomic per arc.  Functors without the member keep the evaluate-both adapter above.
template <typename F, typename = void>
struct cond_guards_apply : std::false_type {};
template <typename F>
struct cond_guards_apply<F, std::void_t<decltype(F::cond_advance_guards_apply)>> : std::integral_constant<bool, F::cond_advance_guards_apply> {};

template <typename Problem, typename Functor>
struct RefFunctorGuardQ {
    typename Problem::data_slice_t *data;
    int iteration;
    const int *offsets;
    static constexpr bool WEIGHTED = false;
    using SrcVal = b200::NoSrc;
    using Evidence = bool;
    using Token = bool;
    using Cand = b200::CandWords<3>;   // dst, src, arc id
    __device__ __forceinline__ SrcVal load_src(int) const { return SrcVal(); }
    __device__ __forceinline__ Evidence probe_load(bool on, SrcVal, int src, int dst, uint32_t eid) const {
        return on && Functor::cond_advance(src, dst, (int)eid, (int)eid - offsets[src], (int)eid, data, iteration);
    }
    __device__ __forceinline__ bool probe_eval(Evidence ok, SrcVal, int, float) const { return ok; }
    __device__ __forceinline__ Cand make_cand(SrcVal, int src, int dst, uint32_t eid, float) const {
        return Cand{{(uint32_t)dst, (uint32_t)src, eid}};
    }
    __device__ __forceinline__ Token claim(const Cand &c) const {
        const int src = (int)c.w[1], dst = (int)c.w[0], eid = (int)c.w[2];
        return Functor::apply_advance(src, dst, eid, eid - offsets[src], eid, data, iteration);
    }
    __device__ __forceinline__ int finish(Token accepted, const Cand &c) const { return accepted ? (int)c.w[0] : -1; }
};

// degree scan of `input` over `offsets`; also mirrors the scan into the graph's
// d_scanned_row_offsets like the reference does (advance.hxx:40).
inline void scan_frontier(b200_workspace *ws, const int *frontier, size_t len, const int *offsets, int *mirror,
                          size_t mirror_capacity) {
    check(b200::reset_counters(ws));
    check(b200::launch_frontier_scan(ws, frontier, (uint32_t)len, reinterpret_cast<const uint32_t *>(offsets)));
    if (mirror && len <= mirror_capacity)
        check(cudaMemcpyAsync(mirror, ws->d_scanned, sizeof(int) * len, cudaMemcpyDeviceToDevice, b200::ws_stream(ws)));
}
}  // namespace detail

template <typename Problem, typename Functor, bool idempotence, bool has_output>
int advance_forward_kernel(std::shared_ptr<Problem> problem, std::shared_ptr<frontier_t<int>> &input,
                           std::shared_ptr<frontier_t<int>> &output, int iteration, standard_context_t &context) {
    const size_t len = input->size();
    if (!len) {
        if (has_output) {
            output->resize(0);
            output->set_hole_free(true);
        }
        return 0;
    }
    graph_device_t &g = *problem->gslice;
    if (b200_ctx_reserve(context.engine(), (int64_t)len) != B200_OK) throw cuda_exception_t(cudaErrorMemoryAllocation);
    b200_workspace *ws = context.workspace();
    const int *in = input->data()->data();
    int *out = has_output ? output->data()->data() : nullptr;
    const unsigned long long cap = has_output ? output->capacity() : 0;
#ifndef B200_ADVANCE_RAW_OUTPUT
    // default: the quad kernel (the raw -1-holed layout and index arrays that are not 16-byte aligned take the
    // arc-wise kernel below; -DB200_ADVANCE_LBS forces it)
#ifndef B200_ADVANCE_LBS
    if ((reinterpret_cast<uintptr_t>(g.d_col_indices.data()) & 15u) == 0) {
        const uint32_t *off = reinterpret_cast<const uint32_t *>(g.d_row_offsets.data());
        detail::check(b200::reset_counters(ws));
        detail::check(b200::launch_quad_scan(ws, in, (uint32_t)len, off));
        const b200::QuadArgs qa = b200::make_quad_args(ws, in, (uint32_t)len, off, g.d_col_indices.data(), nullptr);
        using QOp = typename std::conditional<!idempotence && detail::cond_guards_apply<Functor>::value,
                                              detail::RefFunctorGuardQ<Problem, Functor>,
                                              detail::RefFunctorQ<Problem, Functor, idempotence>>::type;
        QOp qop{problem->d_data_slice.data(), iteration, g.d_row_offsets.data()};
        if (has_output) detail::check((b200::launch_quad_advance<b200::OUT_COMPACT, false>(ws, qa, qop, out, cap)));
        else detail::check((b200::launch_quad_advance<b200::OUT_NONE, false>(ws, qa, qop, nullptr, 0ull)));
        detail::check(b200::read_counters(ws));
        if (!has_output) return 0;
        const size_t produced = (size_t)ws->h_counters[B200_CNT_OUT];
        output->resize(produced);   // prints the reference's overflow message and exits if it does not fit
        output->set_hole_free(true);
        return (int)produced;
    }
#endif
#endif
    detail::scan_frontier(ws, in, len, g.d_row_offsets.data(), g.d_scanned_row_offsets.data(), g.d_scanned_row_offsets.size());
    const b200::LbsArgs a = b200::make_lbs_args(ws, in, (uint32_t)len, reinterpret_cast<const uint32_t *>(g.d_row_offsets.data()),
                                                g.d_col_indices.data());
    detail::RefFunctorOp<Problem, Functor, idempotence> op{problem->d_data_slice.data(), iteration};
    if (!has_output) {
        detail::check(b200::launch_lbs_advance<b200::OUT_NONE, false>(ws, a, op, out, cap));
    } else {
#ifdef B200_ADVANCE_RAW_OUTPUT
        detail::check(b200::launch_lbs_advance<b200::OUT_RAW, false>(ws, a, op, out, cap));
#else
        detail::check(b200::launch_lbs_advance<b200::OUT_COMPACT, false>(ws, a, op, out, cap));
#endif
    }
    detail::check(b200::read_counters(ws));
    if (!has_output) return 0;
    const size_t produced = (size_t)ws->h_counters[B200_CNT_OUT];
    output->resize(produced);   // prints the reference's overflow message and exits if it does not fit
#ifdef B200_ADVANCE_RAW_OUTPUT
    output->set_hole_free(false);
#else
    output->set_hole_free(true);
#endif
    return (int)produced;
}

// dense[v] = cond_sparse_to_dense(v) ? 1 : 0 for every v in `sparse` (advance.hxx:69-84).
template <typename Problem, typename Functor>
void sparse_to_dense_kernel(std::shared_ptr<Problem> problem, std::shared_ptr<frontier_t<int>> &sparse,
                            std::shared_ptr<frontier_t<int>> &dense, int iteration, standard_context_t &context) {
    const int *items = sparse->data()->data();
    int *flags = dense->data()->data();
    dense->set_hole_free(false);   // (per-vertex flags, not a list)
    typename Problem::data_slice_t *data = problem->d_data_slice.data();
    transform([=] __device__(int idx) {
        const int v = items[idx];
        flags[v] = Functor::cond_sparse_to_dense(v, data, iteration) ? 1 : 0;
    }, sparse->size(), context);
}

namespace detail {
template <typename Problem, typename Functor>
struct GenUnvisitedPred {
    const int *items;
    typename Problem::data_slice_t *data;
    int iteration;
    __device__ __forceinline__ bool operator()(uint32_t idx, int &item) const {
        item = items[idx];
        return Functor::cond_gen_unvisited(item, data, iteration);
    }
};

// One thread per unvisited vertex; stops at the first in-neighbour that is in the
// frontier flags and wins apply_advance.  (The reference expands every in-arc of every
// unvisited vertex through LBS, advance.hxx:140-155; results are identical because
// apply_advance can succeed only once per vertex.)
template <typename Problem, typename Functor>
__global__ void pull_list_kernel(int *unvisited, uint32_t count, const int *__restrict__ col_offsets,
                                 const int *__restrict__ row_indices, const int *__restrict__ flags_in, int *flags_out,
                                 typename Problem::data_slice_t *data, int iteration, unsigned long long *arcs) {
    unsigned long long seen = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const int v = unvisited[i];
        if (v < 0) continue;
        const int b = col_offsets[v], e = col_offsets[v + 1];
        for (int k = b; k < e; ++k) {
            const int nbr = row_indices[k];
            ++seen;
            if (flags_in[nbr] && Functor::apply_advance(nbr, v, k, k - b, k, data, iteration)) {
                flags_out[v] = 1;
                unvisited[i] = -1;
                break;
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) seen += __shfl_xor_sync(0xffffffffu, seen, d);
    if ((threadIdx.x & 31) == 0 && seen) atomicAdd(arcs, seen);
}
}  // namespace detail

// compact {item in indices : cond_gen_unvisited(item)} (advance.hxx:86-106); order preserved.
template <typename Problem, typename Functor>
int gen_unvisited_kernel(std::shared_ptr<Problem> problem, std::shared_ptr<frontier_t<int>> &indices,
                         std::shared_ptr<frontier_t<int>> &unvisited, int iteration, standard_context_t &context) {
    const size_t len = indices->size();
    if (!len) {
        unvisited->resize(0);
        return 0;
    }
    if (b200_ctx_reserve(context.engine(), (int64_t)len) != B200_OK) throw cuda_exception_t(cudaErrorMemoryAllocation);
    b200_workspace *ws = context.workspace();
    detail::check(b200::reset_counters(ws));
    detail::GenUnvisitedPred<Problem, Functor> pred{indices->data()->data(), problem->d_data_slice.data(), iteration};
    detail::check(b200::launch_compact(ws, pred, (uint32_t)len, unvisited->data()->data(), unvisited->capacity(),
                                       ws->d_counters + B200_CNT_OUT, ws->d_counters + B200_CNT_OVERFLOW));
    detail::check(b200::read_counters(ws));
    unvisited->resize((size_t)ws->h_counters[B200_CNT_OUT]);
    unvisited->set_hole_free(true);
    return (int)unvisited->size();
}

// Pull step (advance.hxx:108-160): for every v in `unvisited` whose in-neighbourhood
// meets the frontier flags `bitmap`, apply_advance(nbr, v) runs, bitmap_out[v] = 1 and
// the list entry becomes -1.  Returns the number of in-arcs inspected.
template <typename Problem, typename Functor>
int advance_backward_kernel(std::shared_ptr<Problem> problem, std::shared_ptr<frontier_t<int>> &unvisited,
                            std::shared_ptr<frontier_t<int>> &bitmap, std::shared_ptr<frontier_t<int>> &bitmap_out,
                            int iteration, standard_context_t &context) {
    const size_t len = unvisited->size();
    if (!len) return 0;
    unvisited->set_hole_free(false);    // labelled entries become -1 below
    bitmap_out->set_hole_free(false);
    graph_device_t &g = *problem->gslice;
    b200_workspace *ws = context.workspace();
    detail::check(b200::reset_counters(ws));
    const unsigned grid = (unsigned)std::min<size_t>((len + 255) / 256, (size_t)ws->num_sms * 16);
    detail::pull_list_kernel<Problem, Functor><<<grid, 256, 0, context.stream()>>>(
        unvisited->data()->data(), (uint32_t)len, g.d_col_offsets.data(), g.d_row_indices.data(), bitmap->data()->data(),
        bitmap_out->data()->data(), problem->d_data_slice.data(), iteration, ws->d_counters + B200_CNT_ARCS);
    ws->launches++;
    detail::check(cudaGetLastError());
    detail::check(b200::read_counters(ws));
    return (int)ws->h_counters[B200_CNT_ARCS];
}

}  // namespace advance
}  // namespace oprtr
}  // namespace gunrock
