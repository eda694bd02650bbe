// graph_util.cuh -- helpers for building CUDA graphs with conditional nodes by stream capture
// (level_loop.cu, p2p_bfs.cu).
#pragma once
#include <cuda_runtime.h>
#include "engine.cuh"

namespace b200 {

// inside a build function that declares `int st` and a `fail:` label
#define LL_CUDA(call)                                          \
    do {                                                       \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess) {                               \
            st = ::b200::cuda_status(_e);                      \
            goto fail;                                         \
        }                                                      \
    } while (0)

// Ends a capture into `graph` and returns the nodes the next node of `graph` must depend on.
inline int end_capture(cudaStream_t cs, cudaGraphNode_t *deps, size_t *ndeps, size_t max_deps) {
    cudaStreamCaptureStatus cst;
    const cudaGraphNode_t *d = nullptr;
    size_t nd = 0;
    cudaError_t e = cudaStreamGetCaptureInfo(cs, &cst, nullptr, nullptr, &d, &nd);
    if (e == cudaSuccess && nd > max_deps) e = cudaErrorInvalidValue;
    if (e == cudaSuccess) {
        for (size_t i = 0; i < nd; ++i) deps[i] = d[i];
        *ndeps = nd;
    }
    cudaGraph_t same = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(cs, &same);
    if (e != cudaSuccess) return ::b200::cuda_status(e);
    if (e2 != cudaSuccess) return ::b200::cuda_status(e2);
    return B200_OK;
}


}  // namespace b200
