// Level-1 ray queries behind the C ABI: batch equivalents of Scene::intersect(ray,inter),
// Scene::intersect(ray), Scene::shadowRayTest and Scene::occluded (R/src/scene/scene.cpp:21-81).
//
// Launch shape: persistent warps, grid = SM count x resident blocks per SM (persistent_grid_for); every warp runs the
// pooled scheduler of trace_pooled.cuh (64 rays per warp, state in shared memory, stacks in a global scratch area) and
// pulls work from a global counter until the batch is done.  Trees under 512 nodes use a plain per-thread loop.
#include <string>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include "trace_pooled.cuh"
#include "warp_utils.cuh"

namespace wrt {

constexpr int kTraceBlock = 128;
constexpr int kMaxBlocksPerSm = 12;

__device__ __forceinline__ RayIn load_ray(const wrt_ray* rays, size_t i)
{
    const float4* p = reinterpret_cast<const float4*>(rays + i);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    RayIn r;
    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.dx = a.w; r.dy = b.x; r.dz = b.y; r.tmin = b.z; r.tmax = b.w;
    return r;
}

// Scene::shadowRayTest's verdict from a finished traversal (scene.cpp:58-68)
__device__ __forceinline__ bool visible_from(const RayIn& r, int prim, float t, float px, float py, float pz)
{
    if (prim < 0) return true;
    const float ex = (r.ox + r.dx * t) - px, ey = (r.oy + r.dy * t) - py, ez = (r.oz + r.dz * t) - pz;
    return !(ex < -WRT_EPS) && !(ex > WRT_EPS) && !(ey < -WRT_EPS) && !(ey > WRT_EPS) &&
           !(ez < -WRT_EPS) && !(ez > WRT_EPS);
}

struct ClosestSrc {
    const DevSceneView* sc;
    const wrt_ray* rays; int32_t* prim; float* t_out; float* p3; float* n3; int32_t* inside; int32_t* matid;
    static constexpr bool kCanDecide = false;
    __device__ __forceinline__ void target(size_t, float&, float&, float&) const {}
    __device__ __forceinline__ bool decided(const RayIn&, float) const { return false; }
    __device__ __forceinline__ bool load(size_t i, RayIn& r) const { r = load_ray(rays, i); return true; }
    __device__ __forceinline__ void done(size_t i, const RayIn& r, int id, float t) const
    {
        prim[i] = id;
        if (t_out) t_out[i] = t;
        if (p3 || n3 || inside || matid) {
            HitInfo h = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0, 0 };
            if (id >= 0) fill_hit(*sc, id, r, t, h);
            if (p3) { p3[3 * i] = h.px; p3[3 * i + 1] = h.py; p3[3 * i + 2] = h.pz; }
            if (n3) { n3[3 * i] = h.nx; n3[3 * i + 1] = h.ny; n3[3 * i + 2] = h.nz; }
            if (inside) inside[i] = h.inside;
            if (matid) matid[i] = h.matid;
        }
    }
};

template <bool PRUNED>
__global__ void __launch_bounds__(kTraceBlock, WRT_MIN_BLOCKS)
k_trace_closest(DevSceneView sc, const wrt_ray* __restrict__ rays, size_t n, int32_t* __restrict__ prim,
                float* __restrict__ t_out, float* __restrict__ p3, float* __restrict__ n3,
                int32_t* __restrict__ inside, int32_t* __restrict__ matid, unsigned long long* counter, float4* scratch)
{
    ClosestSrc src = { &sc, rays, prim, t_out, p3, n3, inside, matid };
    trace_rays<PRUNED>(sc, src, counter, n, scratch);
}

struct AnySrc {
    const wrt_ray* rays; uint8_t* hit;
    static constexpr bool kCanDecide = false;
    __device__ __forceinline__ void target(size_t, float&, float&, float&) const {}
    __device__ __forceinline__ bool decided(const RayIn&, float) const { return false; }
    __device__ __forceinline__ bool load(size_t i, RayIn& r) const { r = load_ray(rays, i); return true; }
    __device__ __forceinline__ void done(size_t i, const RayIn&, int id, float) const { hit[i] = id >= 0 ? 1 : 0; }
};

template <bool PRUNED>
__global__ void __launch_bounds__(kTraceBlock, WRT_MIN_BLOCKS)
k_trace_any(DevSceneView sc, const wrt_ray* __restrict__ rays, size_t n, uint8_t* __restrict__ hit,
            unsigned long long* counter, float4* scratch)
{
    AnySrc src = { rays, hit };
    trace_rays<PRUNED>(sc, src, counter, n, scratch);
}

// shadowRayTest: ray given, target point given; visible = 1.0f / 0.0f
struct ShadowTestSrc {
    const wrt_ray* rays; const float* target3; float* visible;
    static constexpr bool kCanDecide = true;      // scheduler 4: stop once the verdict can no longer change (occlusion_decided)
    __device__ __forceinline__ void target(size_t i, float& x, float& y, float& z) const { x = target3[3 * i]; y = target3[3 * i + 1]; z = target3[3 * i + 2]; }
    __device__ __forceinline__ bool decided(const RayIn&, float) const { return false; }
    __device__ __forceinline__ bool load(size_t i, RayIn& r) const { r = load_ray(rays, i); return true; }
    __device__ __forceinline__ void done(size_t i, const RayIn& r, int id, float t) const
    {
        visible[i] = visible_from(r, id, t, target3[3 * i], target3[3 * i + 1], target3[3 * i + 2]) ? 1.0f : 0.0f;
    }
};

template <bool PRUNED>
__global__ void __launch_bounds__(kTraceBlock, WRT_MIN_BLOCKS)
k_trace_shadow(DevSceneView sc, const wrt_ray* __restrict__ rays, const float* __restrict__ target3, size_t n,
               float* __restrict__ visible, unsigned long long* counter, float4* scratch)
{
    ShadowTestSrc src = { rays, target3, visible };
    trace_rays<PRUNED>(sc, src, counter, n, scratch);
}

// occluded(p1, dir, p2): builds Ray(p1, dir) (normalising dir) and negates shadowRayTest
struct OccludedSrc {
    const float* q9; uint8_t* occluded;
    float tx, ty, tz;
    static constexpr bool kCanDecide = true;
    __device__ __forceinline__ void target(size_t i, float& x, float& y, float& z) const { x = q9[9 * i + 6]; y = q9[9 * i + 7]; z = q9[9 * i + 8]; }
    __device__ __forceinline__ bool load(size_t i, RayIn& r)
    {
        const float* q = q9 + 9 * i;
        make_ray(q[0], q[1], q[2], q[3], q[4], q[5], r);
        tx = q[6]; ty = q[7]; tz = q[8];
        return true;
    }
    // exact early exit, see ShadowSrc::decided (wavefront_kernels.cuh)
    __device__ __forceinline__ bool decided(const RayIn& r, float best) const
    {
        const float m = 1.5f * WRT_EPS;
        const float ex = (r.ox + r.dx * best) - tx, ey = (r.oy + r.dy * best) - ty, ez = (r.oz + r.dz * best) - tz;
        return (r.dx > 0.f ? ex < -m : (r.dx < 0.f && ex > m)) || (r.dy > 0.f ? ey < -m : (r.dy < 0.f && ey > m)) ||
               (r.dz > 0.f ? ez < -m : (r.dz < 0.f && ez > m));
    }
    __device__ __forceinline__ void done(size_t i, const RayIn& r, int id, float t) const
    {
        const float* q = q9 + 9 * i;
        occluded[i] = visible_from(r, id, t, q[6], q[7], q[8]) ? 0 : 1;
    }
};

template <bool PRUNED>
__global__ void __launch_bounds__(kTraceBlock, WRT_MIN_BLOCKS)
k_trace_occluded(DevSceneView sc, const float* __restrict__ q9, size_t n, uint8_t* __restrict__ occluded,
                 unsigned long long* counter, float4* scratch)
{
    OccludedSrc src = { q9, occluded, 0.f, 0.f, 0.f };
    trace_rays<PRUNED>(sc, src, counter, n, scratch);
}

__global__ void __launch_bounds__(kTraceBlock)
k_count_visits(DevSceneView sc, const wrt_ray* __restrict__ rays, size_t n, unsigned long long* counter,
               unsigned long long* sums /* inner, leaf, tri, sph */)
{
    size_t base;
    unsigned long long a = 0, b = 0, c = 0, d = 0;
    while (next_chunk(counter, n, base)) {
        const size_t i = base + (threadIdx.x & 31);
        if (i >= n) continue;
        const RayIn r = load_ray(rays, i);
        VisitCounters vc = { 0u, 0u, 0u, 0u };
        float t;
        kd_traverse<false, true>(sc, r, t, &vc);
        a += vc.inner; b += vc.leaf; c += vc.tri; d += vc.sph;
    }
    atomicAdd(&sums[0], a); atomicAdd(&sums[1], b); atomicAdd(&sums[2], c); atomicAdd(&sums[3], d);
}

int ensure_trav_scratch(void** ptr, size_t* bytes)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t need = trav_scratch_bytes(sms * kMaxBlocksPerSm, kTraceBlock);
    if (*bytes >= need) return WRT_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *bytes = 0;
    WRT_CUDA(cudaMalloc(ptr, need));
    *bytes = need;
    return WRT_OK;
}

// Grid of a persistent kernel on the CURRENT device: SM count x resident blocks per SM.  Cached per (device, kernel) —
// a process may drive several devices (wrt_init), so the value must not be a function-local static.
static int persistent_grid_uncached(const void* kernel, int block)
{
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // traversal is bound by L1TEX: kernels that keep no state in shared memory get the whole unified cache as L1
    cudaFuncAttributes fa;
    memset(&fa, 0, sizeof fa);
    if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess && fa.sharedSizeBytes == 0)
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    // experiment knobs: fewer resident blocks / an explicit shared-memory carve-out trade occupancy for L1
    if (fa.sharedSizeBytes != 0) {
        if (const char* e = getenv("WRT_SMEM_CARVEOUT")) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
    if (const char* e = getenv("WRT_BLOCKS_PER_SM")) { const int cap = atoi(e); if (cap > 0 && per_sm > cap) per_sm = cap; }
    if (per_sm < 1) per_sm = 1;
    if (per_sm > kMaxBlocksPerSm) per_sm = kMaxBlocksPerSm;   // the traversal scratch is sized for this many
    return sms * per_sm;
}

int persistent_grid_for(const void* kernel, int block)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(dev, kernel);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const int g = persistent_grid_uncached(kernel, block);
    cache[key] = g;
    return g;
}

// Work counter + traversal-stack scratch of one caller stream.  The device-pointer entry points take a stream; two calls on
// different streams of one scene must not share the counter (one zeroes it while the other fetches work) nor the [depth][slot]
// stack scratch, so every stream seen gets its own pair (created on first use, kept for the life of the scene).
int trace_ctx_for(wrt_scene* sc, cudaStream_t st, unsigned long long** counter, float4** scratch)
{
    for (int i = 0; i < sc->n_trace_ctx; i++)
        if (sc->trace_ctx[i].stream == st) { *counter = sc->trace_ctx[i].counter; *scratch = (float4*)sc->trace_ctx[i].scratch; return WRT_OK; }
    if (sc->n_trace_ctx >= WRT_MAX_TRACE_STREAMS) {
        set_error("wrt_trace_*_dev: more than 16 distinct streams used on one scene"); return WRT_ERR_INVALID;
    }
    wrt_trace_ctx c; memset(&c, 0, sizeof c);
    c.stream = st;
    cudaError_t e = cudaMalloc((void**)&c.counter, 16 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(c.counter, 0, 16 * sizeof(unsigned long long));
    if (e != cudaSuccess) { cudaFree(c.counter); cudaGetLastError(); return cuda_fail(e, "trace context"); }
    const int rc = ensure_trav_scratch(&c.scratch, &c.scratch_bytes);
    if (rc) { cudaFree(c.counter); cudaGetLastError(); return rc; }
    sc->trace_ctx[sc->n_trace_ctx++] = c;
    *counter = c.counter; *scratch = (float4*)c.scratch;
    return WRT_OK;
}

// The pooled scheduler keeps a work item's index in 32 bits (PoolSmem::e.w): one call traces fewer than 2^31 rays (64 GB of wrt_ray).
static int check_batch(size_t n, const char* who)
{
    if (n > (size_t)0x7fffffff) { set_error(std::string(who) + ": at most 2^31 - 1 rays per call (split the batch)"); return WRT_ERR_INVALID; }
    return WRT_OK;
}

// d_counters layout: [0] work counter, [8..11] visit sums
static int launch_closest(wrt_scene* sc, const wrt_ray* d_rays, size_t n, int32_t* d_prim, float* d_t,
                          float* d_p, float* d_n, int32_t* d_inside, int32_t* d_matid, cudaStream_t st)
{
    NvtxRange nvtx_range("wrt: trace closest");
    if (n == 0) return WRT_OK;
    { int rc = check_batch(n, "wrt_trace_closest"); if (rc) return rc; }
    unsigned long long* ctr; float4* scr;
    { int rc = trace_ctx_for(sc, st, &ctr, &scr); if (rc) return rc; }
    WRT_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
    if (sc->traversal_mode == WRT_TRAVERSE_PRUNED) {
        const int grid = persistent_grid_for((const void*)k_trace_closest<true>, kTraceBlock);
        k_trace_closest<true><<<grid, kTraceBlock, 0, st>>>(sc->view, d_rays, n, d_prim, d_t, d_p, d_n, d_inside, d_matid, ctr, scr);
    } else {
        const int grid = persistent_grid_for((const void*)k_trace_closest<false>, kTraceBlock);
        k_trace_closest<false><<<grid, kTraceBlock, 0, st>>>(sc->view, d_rays, n, d_prim, d_t, d_p, d_n, d_inside, d_matid, ctr, scr);
    }
    WRT_CUDA(cudaGetLastError());
    sc->stats.closest_rays += n;
    sc->stats.kernel_launches += 1;
    return WRT_OK;
}

static int launch_occluded(wrt_scene* sc, const float* d_q9, size_t n, uint8_t* d_occ, cudaStream_t st)
{
    NvtxRange nvtx_range("wrt: trace occluded");
    if (n == 0) return WRT_OK;
    { int rc = check_batch(n, "wrt_trace_occluded"); if (rc) return rc; }
    unsigned long long* ctr; float4* scr;
    { int rc = trace_ctx_for(sc, st, &ctr, &scr); if (rc) return rc; }
    WRT_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
    if (sc->traversal_mode == WRT_TRAVERSE_PRUNED) {
        const int grid = persistent_grid_for((const void*)k_trace_occluded<true>, kTraceBlock);
        k_trace_occluded<true><<<grid, kTraceBlock, 0, st>>>(sc->view, d_q9, n, d_occ, ctr, scr);
    } else {
        const int grid = persistent_grid_for((const void*)k_trace_occluded<false>, kTraceBlock);
        k_trace_occluded<false><<<grid, kTraceBlock, 0, st>>>(sc->view, d_q9, n, d_occ, ctr, scr);
    }
    WRT_CUDA(cudaGetLastError());
    sc->stats.shadow_rays += n;
    sc->stats.kernel_launches += 1;
    return WRT_OK;
}

}  // namespace wrt

using namespace wrt;

#define CHECK_SCENE(sc) do { if (!(sc)) { set_error("null scene"); return WRT_ERR_INVALID; } \
    WRT_CUDA(cudaSetDevice((sc)->device)); \
    { int rc__ = ensure_trav_scratch(&(sc)->d_trav_scratch, &(sc)->trav_scratch_bytes); if (rc__) return rc__; } } while (0)

extern "C" {

int wrt_trace_closest_full(wrt_scene* sc, const wrt_ray* rays, size_t n, int32_t* prim, float* t,
                           float* p3, float* n3, int32_t* inside, int32_t* matid)
{
    CHECK_SCENE(sc);
    if (n == 0) return WRT_OK;
    if (!rays || !prim) { set_error("wrt_trace_closest: null buffer"); return WRT_ERR_INVALID; }
    // scratch out: prim(4) t(4) p(12) n(12) inside(4) matid(4) per ray
    int rc = ensure_scratch(sc, n * sizeof(wrt_ray), n * 40);
    if (rc) return rc;
    char* o = (char*)sc->d_scratch_out;
    int32_t* d_prim = (int32_t*)o; float* d_t = (float*)(o + 4 * n);
    float* d_p = p3 ? (float*)(o + 8 * n) : nullptr; float* d_n = n3 ? (float*)(o + 20 * n) : nullptr;
    int32_t* d_in = inside ? (int32_t*)(o + 32 * n) : nullptr; int32_t* d_m = matid ? (int32_t*)(o + 36 * n) : nullptr;
    cudaStream_t st = sc->stream;
    WRT_CUDA(cudaMemcpyAsync(sc->d_scratch_in, rays, n * sizeof(wrt_ray), cudaMemcpyHostToDevice, st));
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    rc = launch_closest(sc, (const wrt_ray*)sc->d_scratch_in, n, d_prim, d_t, d_p, d_n, d_in, d_m, st);
    if (rc) return rc;
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaMemcpyAsync(prim, d_prim, 4 * n, cudaMemcpyDeviceToHost, st));
    if (t) WRT_CUDA(cudaMemcpyAsync(t, d_t, 4 * n, cudaMemcpyDeviceToHost, st));
    if (p3) WRT_CUDA(cudaMemcpyAsync(p3, d_p, 12 * n, cudaMemcpyDeviceToHost, st));
    if (n3) WRT_CUDA(cudaMemcpyAsync(n3, d_n, 12 * n, cudaMemcpyDeviceToHost, st));
    if (inside) WRT_CUDA(cudaMemcpyAsync(inside, d_in, 4 * n, cudaMemcpyDeviceToHost, st));
    if (matid) WRT_CUDA(cudaMemcpyAsync(matid, d_m, 4 * n, cudaMemcpyDeviceToHost, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    sc->stats.last_trace_ms = ms;
    return WRT_OK;
}

int wrt_trace_closest(wrt_scene* sc, const wrt_ray* rays, size_t n, int32_t* prim, float* t)
{
    return wrt_trace_closest_full(sc, rays, n, prim, t, nullptr, nullptr, nullptr, nullptr);
}

int wrt_trace_any(wrt_scene* sc, const wrt_ray* rays, size_t n, uint8_t* hit)
{
    CHECK_SCENE(sc);
    if (n == 0) return WRT_OK;
    if (!rays || !hit) { set_error("wrt_trace_any: null buffer"); return WRT_ERR_INVALID; }
    { int rc0 = check_batch(n, "wrt_trace_any"); if (rc0) return rc0; }
    int rc = ensure_scratch(sc, n * sizeof(wrt_ray), n);
    if (rc) return rc;
    cudaStream_t st = sc->stream;
    WRT_CUDA(cudaMemcpyAsync(sc->d_scratch_in, rays, n * sizeof(wrt_ray), cudaMemcpyHostToDevice, st));
    WRT_CUDA(cudaMemsetAsync(sc->d_counters, 0, sizeof(unsigned long long), st));
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    if (sc->traversal_mode == WRT_TRAVERSE_PRUNED) {
        const int grid = persistent_grid_for((const void*)k_trace_any<true>, kTraceBlock);
        k_trace_any<true><<<grid, kTraceBlock, 0, st>>>(sc->view, (const wrt_ray*)sc->d_scratch_in, n, (uint8_t*)sc->d_scratch_out, sc->d_counters, (float4*)sc->d_trav_scratch);
    } else {
        const int grid = persistent_grid_for((const void*)k_trace_any<false>, kTraceBlock);
        k_trace_any<false><<<grid, kTraceBlock, 0, st>>>(sc->view, (const wrt_ray*)sc->d_scratch_in, n, (uint8_t*)sc->d_scratch_out, sc->d_counters, (float4*)sc->d_trav_scratch);
    }
    WRT_CUDA(cudaGetLastError());
    sc->stats.closest_rays += n; sc->stats.kernel_launches += 1;
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaMemcpyAsync(hit, sc->d_scratch_out, n, cudaMemcpyDeviceToHost, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1)); sc->stats.last_trace_ms = ms;
    return WRT_OK;
}

int wrt_trace_shadow(wrt_scene* sc, const wrt_ray* rays, const float* target3, size_t n, float* visible)
{
    CHECK_SCENE(sc);
    if (n == 0) return WRT_OK;
    if (!rays || !target3 || !visible) { set_error("wrt_trace_shadow: null buffer"); return WRT_ERR_INVALID; }
    { int rc0 = check_batch(n, "wrt_trace_shadow"); if (rc0) return rc0; }
    int rc = ensure_scratch(sc, n * (sizeof(wrt_ray) + 16), n * 4);  // rays, then targets (16-byte aligned start)
    if (rc) return rc;
    cudaStream_t st = sc->stream;
    char* in = (char*)sc->d_scratch_in;
    float* d_target = (float*)(in + n * sizeof(wrt_ray));
    WRT_CUDA(cudaMemcpyAsync(in, rays, n * sizeof(wrt_ray), cudaMemcpyHostToDevice, st));
    WRT_CUDA(cudaMemcpyAsync(d_target, target3, n * 12, cudaMemcpyHostToDevice, st));
    WRT_CUDA(cudaMemsetAsync(sc->d_counters, 0, sizeof(unsigned long long), st));
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    if (sc->traversal_mode == WRT_TRAVERSE_PRUNED) {
        const int grid = persistent_grid_for((const void*)k_trace_shadow<true>, kTraceBlock);
        k_trace_shadow<true><<<grid, kTraceBlock, 0, st>>>(sc->view, (const wrt_ray*)in, d_target, n, (float*)sc->d_scratch_out, sc->d_counters, (float4*)sc->d_trav_scratch);
    } else {
        const int grid = persistent_grid_for((const void*)k_trace_shadow<false>, kTraceBlock);
        k_trace_shadow<false><<<grid, kTraceBlock, 0, st>>>(sc->view, (const wrt_ray*)in, d_target, n, (float*)sc->d_scratch_out, sc->d_counters, (float4*)sc->d_trav_scratch);
    }
    WRT_CUDA(cudaGetLastError());
    sc->stats.shadow_rays += n; sc->stats.kernel_launches += 1;
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaMemcpyAsync(visible, sc->d_scratch_out, n * 4, cudaMemcpyDeviceToHost, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1)); sc->stats.last_trace_ms = ms;
    return WRT_OK;
}

int wrt_trace_occluded(wrt_scene* sc, const float* q9, size_t n, uint8_t* occluded)
{
    CHECK_SCENE(sc);
    if (n == 0) return WRT_OK;
    if (!q9 || !occluded) { set_error("wrt_trace_occluded: null buffer"); return WRT_ERR_INVALID; }
    int rc = ensure_scratch(sc, n * 36, n);
    if (rc) return rc;
    cudaStream_t st = sc->stream;
    WRT_CUDA(cudaMemcpyAsync(sc->d_scratch_in, q9, n * 36, cudaMemcpyHostToDevice, st));
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    rc = launch_occluded(sc, (const float*)sc->d_scratch_in, n, (uint8_t*)sc->d_scratch_out, st);
    if (rc) return rc;
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaMemcpyAsync(occluded, sc->d_scratch_out, n, cudaMemcpyDeviceToHost, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1)); sc->stats.last_trace_ms = ms;
    return WRT_OK;
}

int wrt_trace_closest_dev(wrt_scene* sc, const wrt_ray* d_rays, size_t n, int32_t* d_prim, float* d_t, void* stream)
{
    CHECK_SCENE(sc);
    if (n == 0) return WRT_OK;
    if (!d_rays || !d_prim || ((uintptr_t)d_rays & 15)) { set_error("wrt_trace_closest_dev: null or misaligned (16 B) device buffer"); return WRT_ERR_INVALID; }
    return launch_closest(sc, d_rays, n, d_prim, d_t, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int wrt_trace_occluded_dev(wrt_scene* sc, const float* d_q9, size_t n, uint8_t* d_occ, void* stream)
{
    CHECK_SCENE(sc);
    if (n == 0) return WRT_OK;
    if (!d_q9 || !d_occ) { set_error("wrt_trace_occluded_dev: null device buffer"); return WRT_ERR_INVALID; }
    return launch_occluded(sc, d_q9, n, d_occ, (cudaStream_t)stream);
}

int wrt_trace_count_visits(wrt_scene* sc, const wrt_ray* rays, size_t n)
{
    CHECK_SCENE(sc);
    sc->stats.inner_visits = sc->stats.leaf_visits = sc->stats.tri_tests = sc->stats.sphere_tests = 0;
    if (n == 0) return WRT_OK;
    if (!rays) { set_error("wrt_trace_count_visits: null buffer"); return WRT_ERR_INVALID; }
    int rc = ensure_scratch(sc, n * sizeof(wrt_ray), 0);
    if (rc) return rc;
    cudaStream_t st = sc->stream;
    WRT_CUDA(cudaMemcpyAsync(sc->d_scratch_in, rays, n * sizeof(wrt_ray), cudaMemcpyHostToDevice, st));
    WRT_CUDA(cudaMemsetAsync(sc->d_counters, 0, 16 * sizeof(unsigned long long), st));
    const int grid = persistent_grid_for((const void*)k_count_visits, kTraceBlock);
    k_count_visits<<<grid, kTraceBlock, 0, st>>>(sc->view, (const wrt_ray*)sc->d_scratch_in, n, sc->d_counters, sc->d_counters + 8);
    WRT_CUDA(cudaGetLastError());
    sc->stats.kernel_launches += 1;
    unsigned long long h[4];
    WRT_CUDA(cudaMemcpyAsync(h, sc->d_counters + 8, sizeof h, cudaMemcpyDeviceToHost, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    sc->stats.inner_visits = h[0]; sc->stats.leaf_visits = h[1]; sc->stats.tri_tests = h[2]; sc->stats.sphere_tests = h[3];
    return WRT_OK;
}

}  // extern "C"
