// Kernels shared by the PT and BDPT wavefronts: closest-hit "extend" over a queue of path slots and
// the shadow/connection kernel (Scene::occluded + film accumulation), plus their counting-mode
// variants (EXACT traversal with the reference-semantics visit counters).
#pragma once
#include "wavefront.h"
#include "warp_utils.cuh"
#include "trace_pooled.cuh"

namespace wrt {

constexpr int kBlock = 128;

__device__ __forceinline__ RayIn pool_load_ray(const PathPool& pool, uint32_t slot)
{
    const float4* p = reinterpret_cast<const float4*>(pool.ray + slot);
    const float4 a = p[0], b = p[1];
    RayIn r;
    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.dx = a.w; r.dy = b.x; r.dz = b.y; r.tmin = b.z; r.tmax = b.w;
    return r;
}

__device__ __forceinline__ void film_add(float* film, uint32_t pixel, V3 c, float scale)
{
    if (pixel == 0xffffffffu) return;     // off-film sample (ImageFilm::addColor bounds check, film.cpp:6-7)
    atomicAdd(&film[3 * (size_t)pixel + 0], c.x * scale);
    atomicAdd(&film[3 * (size_t)pixel + 1], c.y * scale);
    atomicAdd(&film[3 * (size_t)pixel + 2], c.z * scale);
}

struct ExtendSrc {
    PathPool pool; const uint32_t* queue;
    __device__ __forceinline__ bool load(size_t e, RayIn& r) const { r = pool_load_ray(pool, queue[e]); return true; }
    __device__ __forceinline__ void done(size_t e, const RayIn&, int prim, float t) const
    {
        const uint32_t slot = queue[e];
        pool.hit_prim[slot] = prim;
        pool.hit_t[slot] = t;
    }
};

template <bool PRUNED>
__global__ void __launch_bounds__(kBlock)
k_pt_extend(DevSceneView sc, PathPool pool, const uint32_t* __restrict__ queue, size_t n, unsigned long long* counters,
            float4* scratch)
{
    ExtendSrc src = { pool, queue };
    trace_rays<PRUNED>(sc, src, &counters[WF_WORK], n, scratch);
}

// Counting-mode variants: EXACT traversal with the reference-semantics visit counters.
static __global__ void __launch_bounds__(kBlock)
k_pt_extend_count(DevSceneView sc, PathPool pool, const uint32_t* __restrict__ queue, size_t n, unsigned long long* counters)
{
    size_t base;
    unsigned long long a = 0, b = 0, c = 0, d = 0;
    while (next_chunk(&counters[WF_WORK], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        if (e >= n) continue;
        const uint32_t slot = queue[e];
        const RayIn r = pool_load_ray(pool, slot);
        VisitCounters vc = { 0u, 0u, 0u, 0u };
        float t;
        const int prim = kd_traverse<false, true>(sc, r, t, &vc);
        pool.hit_prim[slot] = prim;
        pool.hit_t[slot] = t;
        a += vc.inner; b += vc.leaf; c += vc.tri; d += vc.sph;
    }
    atomicAdd(&counters[WF_VISITS + 0], a); atomicAdd(&counters[WF_VISITS + 1], b);
    atomicAdd(&counters[WF_VISITS + 2], c); atomicAdd(&counters[WF_VISITS + 3], d);
}

static __global__ void __launch_bounds__(kBlock)
k_pt_shadow_count(DevSceneView sc, ShadowQueue sq, float* __restrict__ film, float scale, unsigned long long* counters)
{
    const size_t n = (size_t)counters[WF_SHADOW_COUNT];
    size_t base;
    unsigned long long a = 0, b = 0, c = 0, d = 0;
    while (next_chunk(&counters[WF_WORK3], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        if (e >= n) continue;
        const float4 qa = sq.a[e], qb = sq.b[e], qc = sq.c[e];
        RayIn r;
        make_ray(qa.x, qa.y, qa.z, qb.x, qb.y, qb.z, r);
        VisitCounters vc = { 0u, 0u, 0u, 0u };
        float t;
        const int prim = kd_traverse<false, true>(sc, r, t, &vc);
        a += vc.inner; b += vc.leaf; c += vc.tri; d += vc.sph;
        bool vis = prim < 0;
        if (!vis) {
            const float ex = (r.ox + r.dx * t) - qc.x, ey = (r.oy + r.dy * t) - qc.y, ez = (r.oz + r.dz * t) - qc.z;
            vis = !(ex < -WRT_EPS) && !(ex > WRT_EPS) && !(ey < -WRT_EPS) && !(ey > WRT_EPS) && !(ez < -WRT_EPS) && !(ez > WRT_EPS);
        }
        if (vis) film_add(film, sq.pixel[e], v3(qa.w, qb.w, qc.w), scale);
    }
    atomicAdd(&counters[WF_VISITS + 0], a); atomicAdd(&counters[WF_VISITS + 1], b);
    atomicAdd(&counters[WF_VISITS + 2], c); atomicAdd(&counters[WF_VISITS + 3], d);
}

struct ShadowSrc {
    ShadowQueue sq; float* film; float scale;
    __device__ __forceinline__ bool load(size_t e, RayIn& r) const
    {
        const float4 a = sq.a[e], b = sq.b[e];
        make_ray(a.x, a.y, a.z, b.x, b.y, b.z, r);
        return true;
    }
    __device__ __forceinline__ void done(size_t e, const RayIn& r, int prim, float t) const
    {
        const float4 c = sq.c[e];
        bool vis = prim < 0;
        if (!vis) {   // Scene::shadowRayTest: the hit point equals the target within EPS per component
            const float ex = (r.ox + r.dx * t) - c.x, ey = (r.oy + r.dy * t) - c.y, ez = (r.oz + r.dz * t) - c.z;
            vis = !(ex < -WRT_EPS) && !(ex > WRT_EPS) && !(ey < -WRT_EPS) && !(ey > WRT_EPS) && !(ez < -WRT_EPS) && !(ez > WRT_EPS);
        }
        if (vis) film_add(film, sq.pixel[e], v3(sq.a[e].w, sq.b[e].w, c.w), scale);
    }
};

template <bool PRUNED>
__global__ void __launch_bounds__(kBlock)
k_pt_shadow(DevSceneView sc, ShadowQueue sq, float* __restrict__ film, float scale, unsigned long long* counters,
            float4* scratch)
{
    const size_t n = (size_t)counters[WF_SHADOW_COUNT];
    ShadowSrc src = { sq, film, scale };
    trace_rays<PRUNED>(sc, src, &counters[WF_WORK3], n, scratch);
}

}  // namespace wrt
