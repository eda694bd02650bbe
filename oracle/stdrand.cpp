// stdrand.cpp -- the C++ standard library's own std::mt19937 + std::uniform_int_distribution<int>, behind a C entry point.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  The reference's colouring problem draws its vertex hashes
// with mgpu::fill_random(0, prime, n, false, ctx) (externals/moderngpu/src/moderngpu/memory.hxx:112-129; used by
// gunrock/src/coloring/coloring_problem.hxx:44,50): ONE process-wide default-seeded std::mt19937 pushed through
// std::uniform_int_distribution<int>(a, b).  The distribution's algorithm is the standard library's business (libstdc++
// changed it between releases), so the oracle does not restate it: it calls the library the product is built against.
#include <random>

extern "C" {

static std::mt19937 &engine() {
    static std::mt19937 e;
    return e;
}

__attribute__((visibility("default"))) void orc_stdrand_reset(void) { engine() = std::mt19937(); }

// out[i] = the next n draws of uniform_int_distribution<int>(a, b) from the shared engine (a fresh distribution
// object per call, as fill_random constructs one per call)
__attribute__((visibility("default"))) void orc_stdrand_uniform(long long n, int a, int b, int *out) {
    std::uniform_int_distribution<int> d(a, b);
    for (long long i = 0; i < n; ++i) out[i] = d(engine());
}

}  // extern "C"
