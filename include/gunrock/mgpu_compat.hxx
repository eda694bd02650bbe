// mgpu_compat.hxx -- the small slice of the moderngpu 2.0 host API that mini-gunrock
// client code touches (standard_context_t, mem_t, to_mem/from_mem/fill/transform ...),
// re-implemented on top of the B200 engine context.
//
// The reference builds on externals/moderngpu (context.hxx:103-219 standard_context_t,
// memory.hxx:10-97 mem_t/to_mem/from_mem/fill/fill_function, transform.hxx:94-102
// transform).  Nothing of mgpu's kernels is used here: standard_context_t owns a
// b200_ctx (include/b200_frontier.h) whose workspace the operator headers launch into.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <exception>
#include <functional>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "b200/workspace.h"
#include "b200_frontier.h"

namespace mgpu {

enum memory_space_t { memory_space_device = 0, memory_space_host = 1 };
enum { warp_size = 32 };

struct cuda_exception_t : std::exception {
    cudaError_t result;
    explicit cuda_exception_t(cudaError_t r) : result(r) {}
    const char *what() const noexcept override { return cudaGetErrorString(result); }
};
inline void throw_on_error(cudaError_t e) {
    if (e != cudaSuccess) throw cuda_exception_t(e);
}

struct context_t {
    virtual ~context_t() {}
    virtual cudaStream_t stream() = 0;
    virtual void synchronize() = 0;
    virtual void *alloc(size_t bytes, memory_space_t space) = 0;
    virtual void free(void *p, memory_space_t space) = 0;
    virtual b200_ctx *engine() = 0;
    b200_workspace *workspace() { return b200_ctx_workspace(engine()); }
};

// `standard_context_t context;` (tests/bfs/test_bfs.cu:22).  Everything runs on the
// legacy default stream like the reference (context.hxx:103), so client cudaMemcpy
// calls stay ordered with the operator kernels.
class standard_context_t : public context_t {
    b200_ctx *_engine = nullptr;
    cudaStream_t _stream;
    cudaDeviceProp _props;
    cudaEvent_t _timer[2];

   public:
    explicit standard_context_t(bool print_prop = true, cudaStream_t stream_ = 0)
        : _stream(stream_ ? stream_ : cudaStreamLegacy) {
        int dev = 0;
        throw_on_error(cudaGetDevice(&dev));
        throw_on_error(cudaGetDeviceProperties(&_props, dev));
        const int s = b200_ctx_create(&_engine, dev, (void *)_stream);
        if (s != B200_OK) {
            std::fprintf(stderr, "b200_ctx_create: %s\n", b200_status_string(s));
            throw cuda_exception_t((cudaError_t)b200_last_cuda_error());
        }
        throw_on_error(cudaEventCreate(&_timer[0]));
        throw_on_error(cudaEventCreate(&_timer[1]));
        if (print_prop) std::printf("%s\n", device_prop_string().c_str());
    }
    ~standard_context_t() override {
        cudaEventDestroy(_timer[0]);
        cudaEventDestroy(_timer[1]);
        b200_ctx_destroy(_engine);
    }
    standard_context_t(const standard_context_t &) = delete;
    standard_context_t &operator=(const standard_context_t &) = delete;

    const cudaDeviceProp &props() const { return _props; }
    std::string device_prop_string() const {
        char buf[256];
        std::snprintf(buf, sizeof buf, "%s : %d SMs, %.0f MHz, %.1f GB, sm_%d%d (b200 frontier engine, ABI %d)",
                      _props.name, _props.multiProcessorCount, _props.clockRate / 1000.0,
                      _props.totalGlobalMem / 1073741824.0, _props.major, _props.minor, b200_abi_version());
        return buf;
    }
    cudaStream_t stream() override { return _stream; }
    void synchronize() override { throw_on_error(cudaStreamSynchronize(_stream)); }
    void *alloc(size_t bytes, memory_space_t space) override {
        void *p = nullptr;
        // (+16 bytes on the device: index / value arrays are read as aligned 128-bit quads, the last of which may reach
        // past the final element -- b200_frontier.h, b200_graph)
        if (bytes) throw_on_error(space == memory_space_device ? cudaMalloc(&p, bytes + 16) : cudaMallocHost(&p, bytes));
        return p;
    }
    void free(void *p, memory_space_t space) override {
        if (p) throw_on_error(space == memory_space_device ? cudaFree(p) : cudaFreeHost(p));
    }
    b200_ctx *engine() override { return _engine; }
    void timer_begin() { cudaEventRecord(_timer[0], _stream); }
    double timer_end() {
        cudaEventRecord(_timer[1], _stream);
        cudaEventSynchronize(_timer[1]);
        float ms = 0;
        cudaEventElapsedTime(&ms, _timer[0], _timer[1]);
        return ms / 1e3;
    }
};

// RAII array in device (or pinned host) memory; movable, not copyable (memory.hxx:10-58).
template <typename type_t>
class mem_t {
    context_t *_context = nullptr;
    type_t *_pointer = nullptr;
    size_t _size = 0;
    memory_space_t _space = memory_space_device;

   public:
    mem_t() {}
    mem_t(size_t size, context_t &context, memory_space_t space = memory_space_device)
        : _context(&context), _size(size), _space(space) {
        _pointer = static_cast<type_t *>(context.alloc(sizeof(type_t) * size, space));
    }
    mem_t(const mem_t &) = delete;
    mem_t &operator=(const mem_t &) = delete;
    mem_t(mem_t &&rhs) noexcept { swap(rhs); }
    mem_t &operator=(mem_t &&rhs) noexcept {
        swap(rhs);
        return *this;
    }
    ~mem_t() {
        if (_context && _pointer) {
            try { _context->free(_pointer, _space); } catch (...) {}
        }
    }
    void swap(mem_t &rhs) noexcept {
        std::swap(_context, rhs._context);
        std::swap(_pointer, rhs._pointer);
        std::swap(_size, rhs._size);
        std::swap(_space, rhs._space);
    }
    context_t &context() { return *_context; }
    size_t size() const { return _size; }
    type_t *data() const { return _pointer; }
    memory_space_t space() const { return _space; }
};

// ---- copies (memory.hxx:60-97).  Blocking, like the reference's cudaMemcpy. ----
template <typename type_t>
cudaError_t htod(type_t *dest, const type_t *source, size_t count) {
    return cudaMemcpy(dest, source, sizeof(type_t) * count, cudaMemcpyHostToDevice);
}
template <typename type_t>
cudaError_t htod(type_t *dest, const std::vector<type_t> &source) {
    return htod(dest, source.data(), source.size());
}
template <typename type_t>
cudaError_t dtoh(type_t *dest, const type_t *source, size_t count) {
    return cudaMemcpy(dest, source, sizeof(type_t) * count, cudaMemcpyDeviceToHost);
}
template <typename type_t>
cudaError_t dtoh(std::vector<type_t> &dest, const type_t *source, size_t count) {
    dest.resize(count);
    return dtoh(dest.data(), source, count);
}
template <typename type_t>
cudaError_t dtod(type_t *dest, const type_t *source, size_t count) {
    return cudaMemcpy(dest, source, sizeof(type_t) * count, cudaMemcpyDeviceToDevice);
}

template <typename type_t>
mem_t<type_t> to_mem(const std::vector<type_t> &data, context_t &context) {
    mem_t<type_t> mem(data.size(), context);
    throw_on_error(htod(mem.data(), data));
    return mem;
}
template <typename type_t>
std::vector<type_t> from_mem(const mem_t<type_t> &mem) {
    std::vector<type_t> host;
    throw_on_error(dtoh(host, mem.data(), mem.size()));
    return host;
}

// ---- elementwise launch: transform(f, count, context) calls f(index) on the device ----
#ifdef __CUDACC__
template <typename func_t>
__global__ void b200_transform_kernel(func_t f, size_t count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) f((int)i);
}
template <typename func_t>
void transform(func_t f, size_t count, context_t &context) {
    if (!count) return;
    const size_t want = (count + 255) / 256;
    const unsigned grid = (unsigned)(want < 148u * 16u ? want : 148u * 16u);
    b200_transform_kernel<<<grid, 256, 0, context.stream()>>>(f, count);
    context.workspace()->launches++;
    throw_on_error(cudaGetLastError());
}
template <typename type_t, typename func_t>
mem_t<type_t> fill_function(func_t f, size_t count, context_t &context) {
    mem_t<type_t> mem(count, context);
    type_t *p = mem.data();
    transform([=] __device__(int index) { p[index] = f(index); }, count, context);
    return mem;
}
template <typename type_t>
mem_t<type_t> fill(type_t value, size_t count, context_t &context) {
    mem_t<type_t> mem(count, context);
    type_t *p = mem.data();
    transform([=] __device__(int index) { p[index] = value; }, count, context);
    return mem;
}
template <typename type_t>
__device__ __forceinline__ type_t ldg(const type_t *p) { return __ldg(p); }
#endif

// ---- reduction operator tags (operators.hxx): only their identity matters here ----
template <typename type_t> struct plus_t {
    __host__ __device__ type_t operator()(type_t a, type_t b) const { return a + b; }
};
template <typename type_t> struct maximum_t {
    __host__ __device__ type_t operator()(type_t a, type_t b) const { return a > b ? a : b; }
};
template <typename type_t> struct minimum_t {
    __host__ __device__ type_t operator()(type_t a, type_t b) const { return a < b ? a : b; }
};

}  // namespace mgpu
