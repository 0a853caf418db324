// Host/device portability shims.  The per-ray and per-path logic is written once as
// __host__ __device__ functions; nvcc builds the product kernels from it, and the test-only
// "hostsim" library (tests/hostsim, never part of the product .so) compiles the very same code
// with g++ -ffp-contract=off so its control flow can be checked against the oracle on a box with
// no GPU.  The arithmetic is IEEE single precision without FMA on both sides.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WRT_HD __host__ __device__ __forceinline__
#else
#define WRT_HD inline
#ifndef WRT_HOSTSIM_VECTOR_TYPES
#define WRT_HOSTSIM_VECTOR_TYPES
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = { x, y, z, w }; return r; }
struct alignas(8) float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r = { x, y }; return r; }
#endif
#endif

namespace wrt {

WRT_HD float4 ldg4(const float4* p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
WRT_HD int f2i(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int v; memcpy(&v, &f, 4); return v;
#endif
}
WRT_HD unsigned f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    unsigned v; memcpy(&v, &f, 4); return v;
#endif
}
WRT_HD float i2f(int v)
{
#ifdef __CUDA_ARCH__
    return __int_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

}  // namespace wrt
