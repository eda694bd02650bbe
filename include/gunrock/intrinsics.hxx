// intrinsics.hxx -- device helpers client functors use (gunrock/src/intrinsics.hxx:6-22).
#pragma once

namespace gunrock {
namespace util {

__device__ __forceinline__ int LaneId() {
    int lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    return lane;
}

// atomicMin on a float; returns the previous value.  Non-negative operands (distances)
// take one native integer atomicMin on the bit pattern -- IEEE-754 orders non-negative
// floats like their bits -- instead of the reference's CAS loop; anything else falls back
// to compare-and-swap.
__device__ __forceinline__ float atomicMin(float *addr, float val) {
    int *as_int = reinterpret_cast<int *>(addr);
    if (__float_as_int(val) >= 0) {
        const int cur = *as_int;
        if (cur >= 0) return __int_as_float(::atomicMin(as_int, __float_as_int(val)));
    }
    int seen = *as_int;
    while (val < __int_as_float(seen)) {
        const int prev = ::atomicCAS(as_int, seen, __float_as_int(val));
        if (prev == seen) break;
        seen = prev;
    }
    return __int_as_float(seen);
}

}  // namespace util
}  // namespace gunrock
