// rmat.cuh -- counter-based RMAT pair generator + pair weights (device & host).
// Specification: SURVEY.md 8d.  The CPU checker used by the tests restates the same
// arithmetic independently; tests require bit-identical output.
#pragma once
#include <stdint.h>

namespace b200 {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t rmat_key(uint64_t seed) { return mix64(seed ^ 0x243F6A8885A308D3ull); }
__host__ __device__ __forceinline__ uint64_t weight_key(uint64_t seed) { return mix64(seed ^ 0x13198A2E03707344ull); }

// (a,b,c,d) = (0.57, 0.19, 0.19, 0.05) as 32-bit integer thresholds
constexpr uint32_t RMAT_T_A = 0x91EB851Eu, RMAT_T_AB = 0xC28F5C28u, RMAT_T_ABC = 0xF3333333u;

__host__ __device__ __forceinline__ void rmat_pair(uint64_t key, uint64_t e, int scale, uint32_t &u, uint32_t &v) {
    u = 0; v = 0;
    for (int lvl = 0; lvl < scale; lvl += 2) {
        const uint64_t h = mix64(key + ((e << 6) | (uint64_t)(lvl >> 1)));
        uint32_t r = (uint32_t)(h >> 32);
        for (int k = 0; k < 2 && lvl + k < scale; ++k) {
            const uint32_t ub = r >= RMAT_T_AB;
            const uint32_t vb = (r >= RMAT_T_A && r < RMAT_T_AB) || (r >= RMAT_T_ABC);
            u = (u << 1) | ub;
            v = (v << 1) | vb;
            r = (uint32_t)h;
        }
    }
}

__host__ __device__ __forceinline__ float pair_weight(uint64_t wkey, uint32_t u, uint32_t v) {
    const uint32_t lo = u < v ? u : v, hi = u < v ? v : u;
    const uint64_t h = mix64(wkey ^ (((uint64_t)lo << 32) | hi));
    return (float)(1 + (int)((h >> 40) & 63));
}

}  // namespace b200
