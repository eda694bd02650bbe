// partition.cuh -- the vertex -> (owner rank, local row) map of the multi-GPU traversals.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace b200 {

// ---------------------------------------------------------------------------
// Swizzled-cyclic 1D vertex partition (multi-GPU): P = 2^log_p ranks.  Vertex v is row
// v >> log_p of its owner's CSR (global column ids) and is owned by rank
//     owner(v) = (v & (P-1)) ^ swizzle(v >> log_p),   swizzle(r) = top log_p bits of r * 0x9E3779B1.
// A plain cyclic map (owner = v & (P-1)) is badly unbalanced on un-permuted RMAT: every id bit is 0
// with probability a+b = 0.76 (arc-weighted), so rank 0 would hold 0.76^log_p of the arcs (44 % at
// P = 8, 3.5x its share).  XOR-ing the low bits with a hash of the row keeps the map a bijection
// (n/P rows per rank, row and owner recoverable from v with a multiply and two shifts) and spreads
// the arcs to within 0.2 % of m/P at scale 26.  Every bitmap is indexed
// "rank-major": bit(v) = owner(v) * n_local + (v >> log_p), so a rank's own slice is the contiguous
// word range [me * n_local/32, (me+1) * n_local/32) and an all-gather of the slices yields the
// whole bitmap in place.  P = 1 is the identity.  Python mirror: mini_b200/partition.py.
// ---------------------------------------------------------------------------
struct Partition {
    uint32_t log_p;      // log2(P)
    uint32_t me;         // this rank
    uint32_t n_local;    // vertices per rank (multiple of 32)
    __host__ __device__ __forceinline__ static uint32_t swizzle(uint32_t r, uint32_t log_p) {
        return ((r * 0x9E3779B1u) >> 1) >> (31u - log_p);   // two shifts: log_p = 0 must give 0
    }
    __host__ __device__ __forceinline__ uint32_t row(uint32_t v) const { return v >> log_p; }
    __host__ __device__ __forceinline__ uint32_t owner(uint32_t v) const {
        return (v & ((1u << log_p) - 1u)) ^ swizzle(v >> log_p, log_p);
    }
    __host__ __device__ __forceinline__ uint32_t bit(uint32_t v) const { return owner(v) * n_local + row(v); }
    // global id of local row r of rank `rank` (inverse of (owner, row))
    __host__ __device__ __forceinline__ uint32_t global_id(uint32_t rank, uint32_t r) const {
        return (r << log_p) | (rank ^ swizzle(r, log_p));
    }
};

}  // namespace b200
