// Warp-level helpers for the persistent kernels: dynamic work fetch and ballot/popc queue append.
#pragma once
#include <cuda_runtime.h>

namespace wrt {

// The warp pulls the next 32 consecutive work items from a global counter.  Warp-uniform result.
__device__ __forceinline__ bool next_chunk(unsigned long long* counter, size_t n, size_t& base)
{
    unsigned long long b = 0;
    if ((threadIdx.x & 31) == 0) b = atomicAdd(counter, 32ull);
    b = __shfl_sync(0xffffffffu, b, 0);
    base = (size_t)b;
    return b < n;
}

// Warp-aggregated append: every lane of the (converged) warp calls this; lanes with pred get a
// unique slot in the queue whose size lives at *counter.  One atomic per warp.
__device__ __forceinline__ unsigned long long warp_append(unsigned long long* counter, bool pred)
{
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

}  // namespace wrt
