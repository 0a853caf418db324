// Kernels shared by the PT and BDPT wavefronts: closest-hit "extend" over a queue of path slots and
// the shadow/connection kernel (Scene::occluded + film accumulation), plus their counting-mode
// variants (EXACT traversal with the reference-semantics visit counters).
#pragma once
#include "wavefront.h"
#include "warp_utils.cuh"
#include "trace_pooled.cuh"

namespace wrt {

constexpr int kBlock = 128;

// The path pool, the queues and the shadow queue are streamed once per bounce (11.8 GB against a 126 MB L2); the kd nodes,
// leaf records and traversal stacks are what should stay in L2.  WRT_STREAM_HINTS marks the traversal kernels' pool / queue
// accesses evict-first (ld.global.cs / st.global.cs).
#ifndef WRT_STREAM_HINTS
#define WRT_STREAM_HINTS 0
#endif
#if WRT_STREAM_HINTS
#define WRT_LDS4(p) __ldcs(p)
#define WRT_LDSU(p) __ldcs(p)
#define WRT_STS(p, v) __stcs((p), (v))
#else
#define WRT_LDS4(p) (*(p))
#define WRT_LDSU(p) (*(p))
#define WRT_STS(p, v) (*(p) = (v))
#endif


__device__ __forceinline__ RayIn pool_load_ray(const PathPool& pool, uint32_t slot)
{
    const float4* p = reinterpret_cast<const float4*>(pool.ray + slot);
    const float4 a = WRT_LDS4(p), b = WRT_LDS4(p + 1);
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

// A queue holds two runs: regenerated camera rays, filled from the BACK of the array (work items
// [0, n_gen)), then continuing paths, filled from the front (work items [n_gen, n_gen + n_cont)).  Camera rays of
// consecutive samples are coherent (8x4 pixel tiles), bounce rays are not: keeping the two kinds in separate runs
// keeps warps homogeneous.
__device__ __forceinline__ uint32_t queue_slot(const uint32_t* queue, size_t e, size_t n_gen, size_t cap)
{
    return e < n_gen ? WRT_LDSU(queue + (cap - 1 - e)) : WRT_LDSU(queue + (e - n_gen));
}

struct ExtendSrc {
    PathPool pool; const uint32_t* queue; size_t n_gen, cap;
    static constexpr bool kCanDecide = false;
    __device__ __forceinline__ void target(size_t, float&, float&, float&) const {}
    __device__ __forceinline__ bool decided(const RayIn&, float) const { return false; }
    __device__ __forceinline__ bool load(size_t e, RayIn& r) const { r = pool_load_ray(pool, queue_slot(queue, e, n_gen, cap)); return true; }
    __device__ __forceinline__ void done(size_t e, const RayIn&, int prim, float t) const
    {
        const uint32_t slot = queue_slot(queue, e, n_gen, cap);
        WRT_STS(pool.hit_prim + slot, prim);
        WRT_STS(pool.hit_t + slot, t);
    }
};

// Device-driven iteration (wavefront.h): `ctr` is the counter block, `parity` selects the bank this iteration writes;
// the queue length comes from the bank the previous iteration wrote.
__device__ __forceinline__ unsigned long long* wf_cur(unsigned long long* ctr, int parity) { return ctr + parity * WF_BANK; }
__device__ __forceinline__ const unsigned long long* wf_prev(const unsigned long long* ctr, int parity) { return ctr + (parity ^ 1) * WF_BANK; }

template <bool PRUNED>
__global__ void __launch_bounds__(kBlock, WRT_MIN_BLOCKS)
k_pt_extend(DevSceneView sc, PathPool pool, const uint32_t* __restrict__ queue, unsigned long long* ctr, int parity,
            float4* scratch, size_t cap)
{
    const unsigned long long* prev = wf_prev(ctr, parity);
    const size_t n = wf_queue_n(prev), n_gen = (size_t)prev[WF_GEN_COUNT];
    if (n == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&ctr[WF_TOTAL_CLOSEST], (unsigned long long)n); atomicAdd(&ctr[WF_TOTAL_ITERS], 1ull); }
    ExtendSrc src = { pool, queue, n_gen, cap };
    trace_rays<PRUNED>(sc, src, &wf_cur(ctr, parity)[WF_WORK], n, scratch);
}

// Counting-mode variants: EXACT traversal with the reference-semantics visit counters.
template <bool PRUNED>
__global__ void __launch_bounds__(kBlock)
k_pt_extend_count(DevSceneView sc, PathPool pool, const uint32_t* __restrict__ queue, unsigned long long* ctr, int parity, size_t cap)
{
    const unsigned long long* prev = wf_prev(ctr, parity);
    const size_t n = wf_queue_n(prev), n_gen = (size_t)prev[WF_GEN_COUNT];
    if (n == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&ctr[WF_TOTAL_CLOSEST], (unsigned long long)n); atomicAdd(&ctr[WF_TOTAL_ITERS], 1ull); }
    unsigned long long* cur = wf_cur(ctr, parity);
    size_t base;
    unsigned long long a = 0, b = 0, c = 0, d = 0;
    while (next_chunk(&cur[WF_WORK], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        if (e >= n) continue;
        const uint32_t slot = queue_slot(queue, e, n_gen, cap);
        const RayIn r = pool_load_ray(pool, slot);
        VisitCounters vc = { 0u, 0u, 0u, 0u };
        float t;
        const int prim = kd_traverse<PRUNED, true>(sc, r, t, &vc);
        pool.hit_prim[slot] = prim;
        pool.hit_t[slot] = t;
        a += vc.inner; b += vc.leaf; c += vc.tri; d += vc.sph;
    }
    atomicAdd(&ctr[WF_VISITS + 0], a); atomicAdd(&ctr[WF_VISITS + 1], b);
    atomicAdd(&ctr[WF_VISITS + 2], c); atomicAdd(&ctr[WF_VISITS + 3], d);
}

template <bool PRUNED>
__global__ void __launch_bounds__(kBlock)
k_pt_shadow_count(DevSceneView sc, ShadowQueue sq, float* __restrict__ film, float scale, unsigned long long* ctr, int parity)
{
    unsigned long long* cur = wf_cur(ctr, parity);
    const size_t n = (size_t)cur[WF_SHADOW_COUNT];
    if (n == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctr[WF_TOTAL_SHADOW], (unsigned long long)n);
    size_t base;
    unsigned long long a = 0, b = 0, c = 0, d = 0;
    while (next_chunk(&cur[WF_WORK3], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        if (e >= n) continue;
        const float4 qa = sq.a[e], qb = sq.b[e], qc = sq.c[e];
        RayIn r;
        make_ray(qa.x, qa.y, qa.z, qb.x, qb.y, qb.z, r);
        VisitCounters vc = { 0u, 0u, 0u, 0u };
        float t;
        const int prim = kd_traverse<PRUNED, true>(sc, r, t, &vc);
        a += vc.inner; b += vc.leaf; c += vc.tri; d += vc.sph;
        bool vis = prim < 0;
        if (!vis) {
            const float ex = (r.ox + r.dx * t) - qc.x, ey = (r.oy + r.dy * t) - qc.y, ez = (r.oz + r.dz * t) - qc.z;
            vis = !(ex < -WRT_EPS) && !(ex > WRT_EPS) && !(ey < -WRT_EPS) && !(ey > WRT_EPS) && !(ez < -WRT_EPS) && !(ez > WRT_EPS);
        }
        if (vis) film_add(film, sq.pixel[e], v3(qa.w, qb.w, qc.w), scale);
    }
    atomicAdd(&ctr[WF_VISITS + 0], a); atomicAdd(&ctr[WF_VISITS + 1], b);
    atomicAdd(&ctr[WF_VISITS + 2], c); atomicAdd(&ctr[WF_VISITS + 3], d);
}

struct ShadowSrc {
    ShadowQueue sq; float* film; float scale;
    float tx, ty, tz;     // target point of the query this lane is tracing (scheduler 2: one ray per lane)
    static constexpr bool kCanDecide = true;
    __device__ __forceinline__ void target(size_t e, float& x, float& y, float& z) const { const float4 c = WRT_LDS4(sq.c + e); x = c.x; y = c.y; z = c.z; }
    __device__ __forceinline__ bool load(size_t e, RayIn& r)
    {
        const float4 a = WRT_LDS4(sq.a + e), b = WRT_LDS4(sq.b + e), c = WRT_LDS4(sq.c + e);
        make_ray(a.x, a.y, a.z, b.x, b.y, b.z, r);
        tx = c.x; ty = c.y; tz = c.z;
        return true;
    }
    // Scene::shadowRayTest only asks whether the closest hit point equals the target within EPS per
    // component (scene.cpp:64-67).  `best` only decreases during the traversal, so once the current best hit
    // lies BEFORE the target along some axis by clearly more than EPS, every later (smaller) best does too:
    // the query is occluded whatever the rest of the traversal finds, and the lane can stop.  (A hit beyond
    // the target decides nothing: a closer one may still move the hit point onto the target.)
    __device__ __forceinline__ bool decided(const RayIn& r, float best) const
    {
        const float m = 1.5f * WRT_EPS;
        const float ex = (r.ox + r.dx * best) - tx, ey = (r.oy + r.dy * best) - ty, ez = (r.oz + r.dz * best) - tz;
        return (r.dx > 0.f ? ex < -m : (r.dx < 0.f && ex > m)) || (r.dy > 0.f ? ey < -m : (r.dy < 0.f && ey > m)) ||
               (r.dz > 0.f ? ez < -m : (r.dz < 0.f && ez > m));
    }
    __device__ __forceinline__ void done(size_t e, const RayIn& r, int prim, float t) const
    {
        const float4 c = WRT_LDS4(sq.c + e);
        bool vis = prim < 0;
        if (!vis) {   // Scene::shadowRayTest: the hit point equals the target within EPS per component
            const float ex = (r.ox + r.dx * t) - c.x, ey = (r.oy + r.dy * t) - c.y, ez = (r.oz + r.dz * t) - c.z;
            vis = !(ex < -WRT_EPS) && !(ex > WRT_EPS) && !(ey < -WRT_EPS) && !(ey > WRT_EPS) && !(ez < -WRT_EPS) && !(ez > WRT_EPS);
        }
        if (vis) film_add(film, WRT_LDSU(sq.pixel + e), v3(WRT_LDS4(sq.a + e).w, WRT_LDS4(sq.b + e).w, c.w), scale);
    }
};

template <bool PRUNED>
__global__ void __launch_bounds__(kBlock, WRT_MIN_BLOCKS)
k_pt_shadow(DevSceneView sc, ShadowQueue sq, float* __restrict__ film, float scale, unsigned long long* ctr, int parity,
            float4* scratch)
{
    unsigned long long* cur = wf_cur(ctr, parity);
    const size_t n = (size_t)cur[WF_SHADOW_COUNT];
    if (n == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctr[WF_TOTAL_SHADOW], (unsigned long long)n);
    ShadowSrc src = { sq, film, scale, 0.f, 0.f, 0.f };
    trace_rays<PRUNED>(sc, src, &cur[WF_WORK3], n, scratch);
}

}  // namespace wrt
