// Microbenchmark: per-iteration cost of CUDA-graph conditional nodes on B200 (device-driven loops).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/graph_cond profiles/microbench/graph_cond.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void tiny(int *c) { if (threadIdx.x == 0 && blockIdx.x == 0) c[1]++; }
__global__ void wide(int *c) { if (threadIdx.x == 0 && blockIdx.x == 0) c[2]++; }
__global__ void step(cudaGraphConditionalHandle hw, cudaGraphConditionalHandle hi, int use_if, int *c) {
    const int v = --c[0];
    cudaGraphSetConditional(hw, v > 0);
    if (use_if) cudaGraphSetConditional(hi, v & 1);
}
__global__ void reset(int *c, int n) { c[0] = n; c[1] = 0; c[2] = 0; }

// variant: 0 = while{step}; 1 = while{step + K tiny kernels}; 2 = while{ if/else{tiny} ; step }; 3 = while{ if/else; step; switch }
int run(int variant, int K, int iters, bool widek) {
    int *c; CK(cudaMalloc(&c, 64));
    cudaStream_t cs, s; CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaGraph_t G; CK(cudaGraphCreate(&G, 0));
    cudaGraphConditionalHandle hw, hi, hs;
    CK(cudaGraphConditionalHandleCreate(&hw, G, 1, cudaGraphCondAssignDefault));
    hi = hw; hs = hw;   // a handle must belong to a conditional node: only create the ones a variant uses
    if (variant >= 2) CK(cudaGraphConditionalHandleCreate(&hi, G, 0, cudaGraphCondAssignDefault));
    if (variant == 3) CK(cudaGraphConditionalHandleCreate(&hs, G, 2, cudaGraphCondAssignDefault));
    cudaGraphNode_t nreset; cudaKernelNodeParams kp = {}; void *a0[] = {&c, &iters}; kp.func = (void *)reset; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = a0;
    CK(cudaGraphAddKernelNode(&nreset, G, nullptr, 0, &kp));
    cudaGraphNodeParams wp = {}; wp.type = cudaGraphNodeTypeConditional; wp.conditional.handle = hw; wp.conditional.type = cudaGraphCondTypeWhile; wp.conditional.size = 1;
    cudaGraphNode_t nw; CK(cudaGraphAddNode(&nw, G, &nreset, 1, &wp));
    cudaGraph_t body = wp.conditional.phGraph_out[0];
    cudaGraphNode_t last = nullptr; int have_last = 0;
    int use_if = variant >= 2;
    if (variant >= 2) {
        cudaGraphNodeParams ip = {}; ip.type = cudaGraphNodeTypeConditional; ip.conditional.handle = hi; ip.conditional.type = cudaGraphCondTypeIf; ip.conditional.size = 2;
        cudaGraphNode_t ni; CK(cudaGraphAddNode(&ni, body, nullptr, 0, &ip));
        for (int b = 0; b < 2; ++b) {
            CK(cudaStreamBeginCaptureToGraph(cs, ip.conditional.phGraph_out[b], nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
            for (int k = 0; k < (K > 0 ? K : 1); ++k) { if (widek) wide<<<148, 1024, 0, cs>>>(c); else tiny<<<1, 32, 0, cs>>>(c); }
            cudaGraph_t same; CK(cudaStreamEndCapture(cs, &same));
        }
        last = ni; have_last = 1;
    }
    CK(cudaStreamBeginCaptureToGraph(cs, body, have_last ? &last : nullptr, nullptr, have_last, cudaStreamCaptureModeRelaxed));
    if (variant == 1) for (int k = 0; k < K; ++k) { if (widek) wide<<<148, 1024, 0, cs>>>(c); else tiny<<<1, 32, 0, cs>>>(c); }
    step<<<1, 1, 0, cs>>>(hw, hi, use_if, c);
    cudaStreamCaptureStatus cst; const cudaGraphNode_t *deps; size_t nd;
    CK(cudaStreamGetCaptureInfo(cs, &cst, nullptr, nullptr, &deps, &nd));
    cudaGraphNode_t dep0 = deps[0];
    cudaGraph_t same; CK(cudaStreamEndCapture(cs, &same));
    if (variant == 3) {
        cudaGraphNodeParams sp = {}; sp.type = cudaGraphNodeTypeConditional; sp.conditional.handle = hs; sp.conditional.type = cudaGraphCondTypeSwitch; sp.conditional.size = 2;
        cudaGraphNode_t ns; CK(cudaGraphAddNode(&ns, body, &dep0, 1, &sp));
        for (int b = 0; b < 2; ++b) {
            CK(cudaStreamBeginCaptureToGraph(cs, sp.conditional.phGraph_out[b], nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
            tiny<<<1, 32, 0, cs>>>(c);
            CK(cudaStreamEndCapture(cs, &same));
        }
    }
    cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, G, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) { CK(cudaGraphLaunch(ex, s)); CK(cudaStreamSynchronize(s)); }
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ex, s)); CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    int h[3]; CK(cudaMemcpy(h, c, 12, cudaMemcpyDeviceToHost));
    printf("variant %d K=%d %s: %.2f us / iteration (%d iterations, counters %d %d %d)\n", variant, K, widek ? "wide" : "tiny", best * 1e3f / iters, iters, h[0], h[1], h[2]);
    return 0;
}

// baseline: K tiny kernels back to back in a plain stream / plain graph
int plain(int K, bool widek) {
    int *c; CK(cudaMalloc(&c, 64)); CK(cudaMemset(c, 0, 64));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, s));
        for (int k = 0; k < K; ++k) { if (widek) wide<<<148, 1024, 0, s>>>(c); else tiny<<<1, 32, 0, s>>>(c); }
        CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("stream launches K=%d %s: %.2f us / kernel\n", K, widek ? "wide" : "tiny", best * 1e3f / K);
    cudaGraph_t G; cudaGraphExec_t ex;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
    for (int k = 0; k < K; ++k) { if (widek) wide<<<148, 1024, 0, s>>>(c); else tiny<<<1, 32, 0, s>>>(c); }
    CK(cudaStreamEndCapture(s, &G)); CK(cudaGraphInstantiate(&ex, G, 0));
    best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ex, s)); CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("plain graph K=%d %s: %.2f us / kernel\n", K, widek ? "wide" : "tiny", best * 1e3f / K);
    // host round trip: kernel + 64-byte D2H + stream sync, repeated
    int *h; CK(cudaMallocHost(&h, 64));
    best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0, s));
        for (int k = 0; k < 100; ++k) { tiny<<<1, 32, 0, s>>>(c); CK(cudaMemcpyAsync(h, c, 64, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
        CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("host round trip (kernel + D2H 64 B + sync): %.2f us / iteration\n", best * 1e3f / 100);
    return 0;
}

int main() {
    if (plain(64, false)) return 1;
    if (plain(64, true)) return 1;
    if (run(0, 0, 1000, false)) return 1;
    if (run(1, 1, 1000, false)) return 1;
    if (run(1, 4, 1000, false)) return 1;
    if (run(1, 4, 1000, true)) return 1;
    if (run(1, 8, 1000, false)) return 1;
    if (run(2, 1, 1000, false)) return 1;
    if (run(2, 2, 1000, false)) return 1;
    if (run(3, 1, 1000, false)) return 1;
    return 0;
}
