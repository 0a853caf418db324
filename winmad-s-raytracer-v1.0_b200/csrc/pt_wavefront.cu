// Wavefront path tracer: the level-2 entry point wrt_render_pt, a drop-in for
// SurfaceIntegrator::render + PathIntegrator::raytracing
// (R/src/surfaceIntegrator/surfaceIntegrator.cpp:14-46, pathIntegrator.cpp:29-148).
//
// Organisation: a persistent pool of path slots in HBM.  Each iteration runs three kernels over
// index queues —
//   extend : closest-hit traversal of every live path's ray           (the hot kernel)
//   shade  : one vertex of the bounce loop per live path; emits at most one shadow query;
//            dead paths are REGENERATED in place from the next camera sample, so the extend queue
//            stays full until the image is finished (persistent ray queues);
//   shadow : Scene::occluded for the queued NEE connections, adds unoccluded contributions.
// Queues are compacted with warp ballot/popc appends (one atomic per warp).  The film is a float
// H x W x 3 accumulator updated with atomicAdd (several samples of one pixel can be in flight).
// Every path owns a counter-based RNG keyed on (seed, pixel, global sample index), so the image is
// independent of pool size, scheduling and of how samples are sharded across GPUs.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include "pt_logic.cuh"
#include "whitted_logic.cuh"
#include "wavefront_kernels.cuh"

namespace wrt {

__device__ __forceinline__ void pool_store(const PathPool& pool, uint32_t slot, const RayIn& r, const PathData& pd)
{
    float4* p = reinterpret_cast<float4*>(pool.ray + slot);
    p[0] = make_float4(r.ox, r.oy, r.oz, r.dx);
    p[1] = make_float4(r.dy, r.dz, r.tmin, r.tmax);
    pool.weight_pdf[slot] = make_float4(pd.weight.x, pd.weight.y, pd.weight.z, pd.last_pdf);
    pool.meta[slot] = make_uint4(pd.pixel, pd.rng.key, pd.rng.ctr, (uint32_t)pd.length | ((uint32_t)pd.last_specular << 16));
}

__device__ __forceinline__ void pool_load_data(const PathPool& pool, uint32_t slot, PathData& pd)
{
    const float4 w = pool.weight_pdf[slot];
    const uint4 m = pool.meta[slot];
    pd.weight = v3(w.x, w.y, w.z); pd.last_pdf = w.w;
    pd.pixel = m.x; pd.rng.key = m.y; pd.rng.ctr = m.z;
    pd.length = (int)(m.w & 0xffffu); pd.last_specular = (int)(m.w >> 16);
}

__global__ void __launch_bounds__(kBlock)
k_pt_init(PtParams P, DevCamera cam, PathPool pool, uint32_t* queue, unsigned n0, unsigned long long first_sample, unsigned long long* ctr)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctr[WF_BANK + WF_NEXT_COUNT] = n0; ctr[WF_BANK + WF_GEN_COUNT] = 0; }   // what iteration 0 (bank 0) reads
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n0; s += gridDim.x * blockDim.x) {
        RayIn r; PathData pd;
        pt_generate(P, cam, first_sample + s, r, pd);
        pool_store(pool, s, r, pd);
        queue[s] = s;
    }
}

__global__ void __launch_bounds__(kBlock)
k_pt_shade(DevSceneView sc, PtParams P, DevCamera cam, PathPool pool, const uint32_t* __restrict__ queue_in,
           uint32_t* __restrict__ queue_out, ShadowQueue sq, float* __restrict__ film, unsigned long long* ctr, int parity,
           unsigned long long* next_sample, size_t cap)
{
    const unsigned long long* prev = wf_prev(ctr, parity);
    unsigned long long* counters = wf_cur(ctr, parity);
    const size_t n = wf_queue_n(prev), n_gen_in = (size_t)prev[WF_GEN_COUNT];
    size_t base;
    while (next_chunk(&counters[WF_WORK2], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        const bool valid = e < n;
        uint32_t slot = 0;
        RayIn r; PathData pd; ShadeOut out;
        out.alive = false; out.emit = false; out.shadow = false;
        if (valid) {
            slot = queue_slot(queue_in, e, n_gen_in, cap);
            r = pool_load_ray(pool, slot);
            pool_load_data(pool, slot, pd);
            pt_shade(sc, P, r, pd, pool.hit_prim[slot], pool.hit_t[slot], out);
            if (out.emit) film_add(film, pd.pixel, out.emit_c, P.film_scale);
        }
        // NEE connection -> shadow queue
        const unsigned long long spos = warp_append(&counters[WF_SHADOW_COUNT], valid && out.shadow);
        if (valid && out.shadow) {
            sq.a[spos] = make_float4(out.q[0], out.q[1], out.q[2], out.shadow_c.x);
            sq.b[spos] = make_float4(out.q[3], out.q[4], out.q[5], out.shadow_c.y);
            sq.c[spos] = make_float4(out.q[6], out.q[7], out.q[8], out.shadow_c.z);
            sq.pixel[spos] = pd.pixel;
        }
        // path regeneration: a finished path's slot takes the next camera sample
        const bool alive = valid && out.alive;
        const bool dead = valid && !out.alive;
        const unsigned long long snew = warp_append(next_sample, dead);
        const bool regen = dead && snew < P.total_samples;
        if (regen) pt_generate(P, cam, snew, r, pd);
        const unsigned long long qpos = warp_append(&counters[WF_NEXT_COUNT], alive);
        const unsigned long long gpos = warp_append(&counters[WF_GEN_COUNT], regen);
        if (alive) { pool_store(pool, slot, r, pd); queue_out[qpos] = slot; }                 // continuing paths: front
        if (regen) { pool_store(pool, slot, r, pd); queue_out[cap - 1 - gpos] = slot; }       // new camera rays: back
    }
}

// ---- Whitted (whitted_logic.cuh): same pool and queues; meta.w = dep | pending << 16; weight.x = w -------------------
struct WhittedPending {        // the parked second child of every level of a slot's ray tree: [level][slot]
    float4* o;                 // origin, w
    float4* d;                 // direction, dep (int bits)
};

__device__ __forceinline__ void wh_store(const PathPool& pool, uint32_t slot, const RayIn& r, float w, uint32_t pixel, const Rng& rng,
                                         int dep, int pending)
{
    float4* p = reinterpret_cast<float4*>(pool.ray + slot);
    p[0] = make_float4(r.ox, r.oy, r.oz, r.dx);
    p[1] = make_float4(r.dy, r.dz, r.tmin, r.tmax);
    pool.weight_pdf[slot] = make_float4(w, 0.f, 0.f, 0.f);
    pool.meta[slot] = make_uint4(pixel, rng.key, rng.ctr, (uint32_t)dep | ((uint32_t)pending << 16));
}

__global__ void __launch_bounds__(kBlock)
k_wh_init(PtParams P, DevCamera cam, PathPool pool, uint32_t* queue, unsigned n0, unsigned long long first_sample, unsigned long long* ctr)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctr[WF_BANK + WF_NEXT_COUNT] = n0; ctr[WF_BANK + WF_GEN_COUNT] = 0; }
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n0; s += gridDim.x * blockDim.x) {
        RayIn r; PathData pd;
        pt_generate(P, cam, first_sample + s, r, pd);
        wh_store(pool, s, r, 1.f, pd.pixel, pd.rng, 0, 0);
        queue[s] = s;
    }
}

__global__ void __launch_bounds__(kBlock)
k_wh_shade(DevSceneView sc, PtParams P, DevCamera cam, PathPool pool, const uint32_t* __restrict__ queue_in,
           uint32_t* __restrict__ queue_out, ShadowQueue sq, float* __restrict__ film, unsigned long long* ctr, int parity,
           unsigned long long* next_sample, size_t cap, WhittedPending pend, int levels)
{
    const unsigned long long* prev = wf_prev(ctr, parity);
    unsigned long long* counters = wf_cur(ctr, parity);
    const size_t n = wf_queue_n(prev), n_gen_in = (size_t)prev[WF_GEN_COUNT];
    size_t base;
    while (next_chunk(&counters[WF_WORK2], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        const bool valid = e < n;
        uint32_t slot = 0, pixel = 0;
        RayIn r; Rng rng = { 0u, 0u }; WhittedOut out;
        float w = 0.f; int dep = 0, pending = 0;
        out.emit = false; out.shadow = false; out.reflect.valid = false; out.trans.valid = false;
        if (valid) {
            slot = queue_slot(queue_in, e, n_gen_in, cap);
            r = pool_load_ray(pool, slot);
            const uint4 m = pool.meta[slot];
            w = pool.weight_pdf[slot].x;
            pixel = m.x; rng.key = m.y; rng.ctr = m.z; dep = (int)(m.w & 0xffffu); pending = (int)(m.w >> 16);
            whitted_shade(sc, P.max_depth, r, dep, w, rng, pool.hit_prim[slot], pool.hit_t[slot], out);
            if (out.emit) film_add(film, pixel, out.emit_c, P.film_scale);
        }
        const unsigned long long spos = warp_append(&counters[WF_SHADOW_COUNT], valid && out.shadow);
        if (valid && out.shadow) {
            sq.a[spos] = make_float4(out.q[0], out.q[1], out.q[2], out.shadow_c.x);
            sq.b[spos] = make_float4(out.q[3], out.q[4], out.q[5], out.shadow_c.y);
            sq.c[spos] = make_float4(out.q[6], out.q[7], out.q[8], out.shadow_c.z);
            sq.pixel[spos] = pixel;
        }
        // next ray of this slot: a child of this node, else the most recently parked child, else a new camera sample
        bool alive = false;
        if (valid) {
            if (out.reflect.valid) {
                if (out.trans.valid && pending < levels) {
                    const size_t at = (size_t)pending * cap + slot;
                    const RayIn& q = out.trans.ray;
                    pend.o[at] = make_float4(q.ox, q.oy, q.oz, out.trans.w);
                    pend.d[at] = make_float4(q.dx, q.dy, q.dz, __int_as_float(dep + 1));
                    pending++;
                }
                r = out.reflect.ray; w = out.reflect.w; dep += 1; alive = true;
            } else if (out.trans.valid) {
                r = out.trans.ray; w = out.trans.w; dep += 1; alive = true;
            } else if (pending > 0) {
                pending--;
                const size_t at = (size_t)pending * cap + slot;
                const float4 o = pend.o[at], d = pend.d[at];
                r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.tmin = 0.f; r.tmax = WRT_INF;
                w = o.w; dep = __float_as_int(d.w); alive = true;
            }
        }
        const bool dead = valid && !alive;
        const unsigned long long snew = warp_append(next_sample, dead);
        const bool regen = dead && snew < P.total_samples;
        if (regen) {
            PathData pd;
            pt_generate(P, cam, snew, r, pd);
            w = 1.f; pixel = pd.pixel; rng = pd.rng; dep = 0; pending = 0;
        }
        const unsigned long long qpos = warp_append(&counters[WF_NEXT_COUNT], alive);
        const unsigned long long gpos = warp_append(&counters[WF_GEN_COUNT], regen);
        if (alive) { wh_store(pool, slot, r, w, pixel, rng, dep, pending); queue_out[qpos] = slot; }
        if (regen) { wh_store(pool, slot, r, w, pixel, rng, dep, pending); queue_out[cap - 1 - gpos] = slot; }
    }
}

// ImageFilm::outputImage per pixel: scale, clamp, gamma, 8-bit (film.cpp:44-61, color.h:47-75)
__global__ void k_film_resolve(const float* __restrict__ film, size_t n, float scale, float inv_gamma, uint8_t* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float c = film[i] * scale;
        c = fminf(1.0f, fmaxf(c, 0.0f));
        c = powf(c, inv_gamma);
        out[i] = (unsigned char)((double)c * 255.0);
    }
}

// ---- wavefront state -------------------------------------------------------------------------------
static void wavefront_free(wrt_wavefront* wf);

// Allocation failures leave NO half-built object behind: the slot is published (and its capacity becomes visible to the
// `capacity >= requested` check of the next call) only after every allocation has succeeded.
static int wavefront_alloc(wrt_wavefront* wf, int capacity)
{
    const size_t P = (size_t)capacity;
    WRT_CUDA(cudaMalloc((void**)&wf->pool.ray, P * sizeof(wrt_ray)));
    WRT_CUDA(cudaMalloc((void**)&wf->pool.weight_pdf, P * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&wf->pool.meta, P * sizeof(uint4)));
    WRT_CUDA(cudaMalloc((void**)&wf->pool.hit_prim, P * sizeof(int32_t)));
    WRT_CUDA(cudaMalloc((void**)&wf->pool.hit_t, P * sizeof(float)));
    WRT_CUDA(cudaMalloc((void**)&wf->queue[0], P * sizeof(uint32_t)));
    WRT_CUDA(cudaMalloc((void**)&wf->queue[1], P * sizeof(uint32_t)));
    WRT_CUDA(cudaMalloc((void**)&wf->shadow.a, P * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&wf->shadow.b, P * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&wf->shadow.c, P * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&wf->shadow.pixel, P * sizeof(uint32_t)));
    WRT_CUDA(cudaMalloc((void**)&wf->shadow2.a, P * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&wf->shadow2.b, P * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&wf->shadow2.c, P * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&wf->shadow2.pixel, P * sizeof(uint32_t)));
    WRT_CUDA(cudaMalloc((void**)&wf->counters, WF_COUNTERS * sizeof(unsigned long long)));
    WRT_CUDA(cudaMallocHost((void**)&wf->h_counters, WF_COUNTERS * sizeof(unsigned long long)));
    WRT_CUDA(cudaStreamCreateWithFlags(&wf->stream, cudaStreamNonBlocking));
    { int rc = ensure_trav_scratch(&wf->trav_scratch, &wf->trav_scratch_bytes); if (rc) return rc; }
    { int rc = ensure_trav_scratch(&wf->trav_scratch2, &wf->trav_scratch2_bytes); if (rc) return rc; }
    WRT_CUDA(cudaStreamCreateWithFlags(&wf->shadow_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        WRT_CUDA(cudaEventCreateWithFlags(&wf->shaded_ev[i], cudaEventDisableTiming));
        WRT_CUDA(cudaEventCreateWithFlags(&wf->shadowed_ev[i], cudaEventDisableTiming));
    }
    WRT_CUDA(cudaEventCreateWithFlags(&wf->join_ev, cudaEventDisableTiming));
    WRT_CUDA(cudaEventCreateWithFlags(&wf->poll_ev[0], cudaEventDisableTiming));
    WRT_CUDA(cudaEventCreateWithFlags(&wf->poll_ev[1], cudaEventDisableTiming));
    return WRT_OK;
}

// Experiment knob (off by default): WRT_L2_PERSIST_MB=<n> sets aside n MB of L2 for persisting lines and puts an access-policy
// window over the kd nodes (WRT_L2_PERSIST_WHAT=recs: over the leaf records) on the traversal streams.  Best effort.
static void apply_l2_window(wrt_scene* sc, cudaStream_t st)
{
    const char* e = getenv("WRT_L2_PERSIST_MB");
    const long mb = e ? atol(e) : 0;
    if (mb <= 0) return;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, sc->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, sc->device);
    size_t want = (size_t)mb << 20;
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    if (want == 0 || max_window <= 0) return;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    const char* w = getenv("WRT_L2_PERSIST_WHAT");
    const bool recs = w && w[0] == 'r';
    size_t bytes = recs ? (size_t)sc->n_leaf_recs * 48 : (size_t)sc->view.n_nodes * 32 + 32;
    if (bytes > (size_t)max_window) bytes = (size_t)max_window;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    v.accessPolicyWindow.base_ptr = recs ? sc->d_leaf_recs : sc->d_nodes;
    v.accessPolicyWindow.num_bytes = bytes;
    v.accessPolicyWindow.hitRatio = bytes <= want ? 1.0f : (float)((double)want / (double)bytes);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();
}

int wavefront_get_slot(wrt_scene* sc, int slot, int capacity, wrt_wavefront** out)
{
    wrt_wavefront** where = slot == 0 ? &sc->wf : &sc->wf_extra[slot - 1];
    if (*where && (*where)->capacity >= capacity) { *out = *where; return WRT_OK; }
    if (*where) { wavefront_free(*where); *where = nullptr; }
    wrt_wavefront* wf = new wrt_wavefront();
    memset(wf, 0, sizeof *wf);
    const int rc = wavefront_alloc(wf, capacity);
    if (rc != WRT_OK) {
        wavefront_free(wf);          // frees whatever was allocated; the slot stays empty
        cudaGetLastError();          // an out-of-memory cudaMalloc is not sticky: clear it for the next call
        return rc;
    }
    wf->capacity = capacity;
    apply_l2_window(sc, wf->stream);
    apply_l2_window(sc, wf->shadow_stream);
    *where = wf;
    *out = wf;
    return WRT_OK;
}

int wavefront_get(wrt_scene* sc, int capacity, wrt_wavefront** out) { return wavefront_get_slot(sc, 0, capacity, out); }

// The library-owned device film of the host-buffer entry points (owned by the scene, so growing a pool never frees it).
int wavefront_film(wrt_scene* sc, size_t floats, float** out)
{
    if (sc->film_floats < floats) {
        if (sc->d_film) cudaFree(sc->d_film);
        sc->d_film = nullptr; sc->film_floats = 0;
        WRT_CUDA(cudaMalloc((void**)&sc->d_film, floats * sizeof(float)));
        sc->film_floats = floats;
    }
    *out = sc->d_film;
    return WRT_OK;
}

void bdpt_destroy(wrt_wavefront* wf);

int wavefront_events(wrt_wavefront* wf, int n)
{
    if (wf->n_ev >= n) return WRT_OK;
    cudaEvent_t* ne = new cudaEvent_t[n];
    for (int i = 0; i < wf->n_ev; i++) ne[i] = wf->ev[i];
    for (int i = wf->n_ev; i < n; i++) WRT_CUDA(cudaEventCreate(&ne[i]));
    delete[] wf->ev;
    wf->ev = ne; wf->n_ev = n;
    return WRT_OK;
}

// Events per iteration (kEvPerIter = 5): [0] before extend, [1] after extend, [2] after shade, [3] before shadow, [4] after shadow.
void wavefront_sum_stage_times(wrt_scene* sc, wrt_wavefront* wf, int iters_timed)
{
    double ext = 0, shd = 0, shw = 0;
    for (int i = 0; i < iters_timed; i++) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, wf->ev[kEvPerIter * i], wf->ev[kEvPerIter * i + 1]);
        cudaEventElapsedTime(&b, wf->ev[kEvPerIter * i + 1], wf->ev[kEvPerIter * i + 2]);
        cudaEventElapsedTime(&c, wf->ev[kEvPerIter * i + 3], wf->ev[kEvPerIter * i + 4]);
        ext += a; shd += b; shw += c;
    }
    sc->stats.extend_ms = ext; sc->stats.shade_ms = shd; sc->stats.shadow_ms = shw;
}

static void wavefront_free(wrt_wavefront* wf)
{
    if (!wf) return;
    bdpt_destroy(wf);
    cudaFree(wf->pool.ray); cudaFree(wf->pool.weight_pdf); cudaFree(wf->pool.meta);
    cudaFree(wf->pool.hit_prim); cudaFree(wf->pool.hit_t);
    cudaFree(wf->queue[0]); cudaFree(wf->queue[1]);
    cudaFree(wf->shadow.a); cudaFree(wf->shadow.b); cudaFree(wf->shadow.c); cudaFree(wf->shadow.pixel);
    cudaFree(wf->shadow2.a); cudaFree(wf->shadow2.b); cudaFree(wf->shadow2.c); cudaFree(wf->shadow2.pixel); cudaFree(wf->trav_scratch2);
    if (wf->shadow_stream) cudaStreamDestroy(wf->shadow_stream);
    for (int i = 0; i < 2; i++) { if (wf->shaded_ev[i]) cudaEventDestroy(wf->shaded_ev[i]); if (wf->shadowed_ev[i]) cudaEventDestroy(wf->shadowed_ev[i]); }
    cudaFree(wf->counters); if (wf->h_counters) cudaFreeHost(wf->h_counters); cudaFree(wf->trav_scratch); cudaFree(wf->whitted);
    for (int i = 0; i < wf->n_ev; i++) cudaEventDestroy(wf->ev[i]);
    delete[] wf->ev;
    if (wf->stream) cudaStreamDestroy(wf->stream);
    if (wf->join_ev) cudaEventDestroy(wf->join_ev);
    for (int i = 0; i < 2; i++) if (wf->poll_ev[i]) cudaEventDestroy(wf->poll_ev[i]);
    delete wf;
}

void wavefront_destroy(wrt_scene* sc)
{
    wavefront_free(sc->wf);
    sc->wf = nullptr;
    for (int i = 0; i < 7; i++) { wavefront_free(sc->wf_extra[i]); sc->wf_extra[i] = nullptr; }
}

void fill_camera(const wrt_camera* c, DevCamera& d)
{
    for (int a = 0; a < 3; a++) { d.pos[a] = c->pos[a]; d.forward[a] = c->forward[a]; }
    d.image_plane_dist = c->image_plane_dist; d.x_res = c->x_res; d.y_res = c->y_res;
    memcpy(d.r2w, c->raster_to_world, sizeof d.r2w);
    memcpy(d.w2r, c->world_to_raster, sizeof d.w2r);
}

static int pool_capacity()
{
    const char* e = getenv("WRT_POOL_PATHS");
    long v = e ? atol(e) : (1L << 26);
    if (v < 1024) v = 1024;
    if (v > (1L << 27)) v = 1L << 27;
    return (int)v;
}

static int pt_fill_params(const wrt_pt_params* p, PtParams& P)
{
    if (!p || p->width <= 0 || p->height <= 0 || p->spp <= 0 || p->max_depth < 0) {
        set_error("wrt_render_pt: bad parameters"); return WRT_ERR_INVALID;
    }
    P.width = p->width; P.height = p->height; P.spp = p->spp; P.max_depth = p->max_depth; P.seed = p->seed;
    P.strata = (int)std::sqrt((double)p->spp);
    if (P.strata < 1) P.strata = 1;
    P.sample_first = p->sample_first; P.sample_stride = p->sample_stride > 0 ? p->sample_stride : 1;
    if (P.sample_first < 0 || P.sample_first >= P.spp) { set_error("wrt_render_pt: sample_first out of range"); return WRT_ERR_INVALID; }
    P.local_spp = (P.spp - P.sample_first + P.sample_stride - 1) / P.sample_stride;
    P.film_scale = p->film_scale != 0.f ? p->film_scale : 1.f / (float)p->spp;
    P.total_samples = (unsigned long long)P.width * P.height * (unsigned long long)P.local_spp;
    P.tape_stride = 0;
    return WRT_OK;
}

static int sub_pools()
{
    const char* e = getenv("WRT_SUBPOOLS");
    int k = e ? atoi(e) : 2;
    return std::min(std::max(k, 1), 8);
}

// How the pool is split: K sub-pools, each with its own queues, counters and stream.  The kernels of
// different sub-pools overlap on the GPU, so the tail of one sub-pool's extend launch (a handful of very
// long rays, ~3.5 ms on the 1 M-triangle scene) is covered by the bulk of another's.
struct PtPlan { int k; int cap[8]; };

static void pt_plan(const PtParams& P, PtPlan& plan, bool whitted = false)
{
    unsigned long long total = std::min<unsigned long long>((unsigned long long)pool_capacity(), P.total_samples);
    if (whitted) total = std::min<unsigned long long>(total, 1ull << 23);    // + 32 bytes x (maxTracingDepth + 1) of pending list per slot
    total = std::max<unsigned long long>(total, 1024ull);
    int k = sub_pools();
    if (total < (1ull << 18)) k = 1;
    // Whitted: the pool is capped at 2^23 slots (pending lists), so halving it makes the launches small: one sub-pool measured
    // 1094 Mrays/s against 862 (two) and 737 (four) on torus.scene (profiles/r2_experiments.md)
    if (whitted && !getenv("WRT_SUBPOOLS")) k = 1;
    plan.k = k;
    // Sub-pool sizes.  When the pool holds the whole frame (no slot is ever regenerated: a frame of <= 2^26 samples, e.g. one
    // GPU's share of a sharded render) every sub-pool runs the same short chain of iterations — one large launch of camera
    // rays, then a handful of small, latency-bound ones — and equal sub-pools would walk through it in lock-step, their
    // tails coinciding.  Unequal shares (WRT_SUBPOOL_WEIGHTS, default 65,35 for two) stagger the chains, so one sub-pool's
    // small iterations run beside the other's large one.  With regeneration the queues stay full and equal shares are used.
    double wgt[8]; double sum = 0;
    const bool staggers = P.total_samples <= total;
    for (int j = 0; j < k; j++) wgt[j] = 1.0;
    if (staggers && k == 2) { wgt[0] = 65; wgt[1] = 35; }
    if (const char* e = getenv("WRT_SUBPOOL_WEIGHTS")) {
        int j = 0; const char* q = e;
        while (*q && j < k) { wgt[j++] = std::max(1.0, atof(q)); while (*q && *q != ',') q++; if (*q == ',') q++; }
    }
    for (int j = 0; j < k; j++) sum += wgt[j];
    for (int j = 0; j < k; j++) {
        const unsigned long long share = (unsigned long long)std::ceil((double)total * wgt[j] / sum);
        plan.cap[j] = (int)std::max<unsigned long long>(std::min<unsigned long long>(share, 1ull << 27), 1024ull);
    }
}

// Host-side view of one sub-pool's device-driven loop.  Iterations are enqueued in batches; after every batch the bank
// the last iteration wrote is copied to one of two pinned slots, so the host learns the queue length one batch late
// and never stalls the stream to find out.
struct SubState {
    wrt_wavefront* wf; int index;
    int iter;                 // iterations enqueued so far (parity of the next one = iter & 1)
    int batches;              // batches enqueued so far
    int polled;               // batches whose result has been read
    size_t last_n;            // queue length seen by the most recent poll
    bool done;
    int timed;
};

// Per-sub-pool pending lists of the Whitted integrator: (max_depth + 1) levels x capacity x 32 bytes.
static int whitted_pending(wrt_wavefront* wf, int levels, WhittedPending& out)
{
    const size_t need = (size_t)levels * (size_t)wf->capacity * 2 * sizeof(float4);
    if (wf->whitted_bytes < need) {
        if (wf->whitted) cudaFree(wf->whitted);
        wf->whitted = nullptr; wf->whitted_bytes = 0;
        WRT_CUDA(cudaMalloc(&wf->whitted, need));
        wf->whitted_bytes = need;
    }
    out.o = (float4*)wf->whitted;
    out.d = out.o + (size_t)levels * (size_t)wf->capacity;
    return WRT_OK;
}

int render_pt_device(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* d_film, cudaStream_t st, bool whitted)
{
    NvtxRange nvtx_range("wrt: PT / Whitted wavefront");
    if (!cam) { set_error("wrt_render_pt: null camera"); return WRT_ERR_INVALID; }
    if (sc->view.n_lights <= 0) { set_error("wrt_render_pt: the scene has no light (the reference indexes an empty vector here)"); return WRT_ERR_INVALID; }
    PtParams P;
    int rc = pt_fill_params(p, P);
    if (rc) return rc;
    if (whitted && P.max_depth > 63) {   // the per-slot pending list holds one parked child per level
        set_error("wrt_render_whitted: max_depth > 63 is not supported (the ray tree's pending list has 64 levels)"); return WRT_ERR_INVALID;
    }
    DevCamera dc; fill_camera(cam, dc);
    // diagnostics: replay the reference's random numbers (wrt_debug_set_rng_tape)
    struct TapeGuard {
        bool on = false;
        ~TapeGuard() { if (on) { const float* none = nullptr; cudaMemcpyToSymbol(c_rng_tape, &none, sizeof none); } }
    } tape_guard;
    if (sc->d_rng_tape) {
        const unsigned long long need = (unsigned long long)P.width * P.height * (unsigned long long)P.spp * sc->rng_tape_stride;
        if (need > sc->rng_tape_floats) { set_error("wrt_render_pt: the installed RNG tape is shorter than width*height*spp*stride"); return WRT_ERR_INVALID; }
        const float* tp = sc->d_rng_tape;
        WRT_CUDA(cudaMemcpyToSymbol(c_rng_tape, &tp, sizeof tp));
        tape_guard.on = true;
        P.tape_stride = sc->rng_tape_stride;
    }
    PtPlan plan; pt_plan(P, plan, whitted);
    SubState sub[8];
    const int wh_levels = std::min(P.max_depth + 1, 64);
    WhittedPending wh_pend[8];
    for (int j = 0; j < plan.k; j++) {
        rc = wavefront_get_slot(sc, j, plan.cap[j], &sub[j].wf);
        if (rc) return rc;
        rc = wavefront_events(sub[j].wf, kEvPerIter * 64);
        if (rc) return rc;
        if (whitted) { rc = whitted_pending(sub[j].wf, wh_levels, wh_pend[j]); if (rc) return rc; }
    }
    const bool pruned = sc->traversal_mode == WRT_TRAVERSE_PRUNED;
    const bool counting = sc->counting != 0;
    const bool count_pruned = sc->counting == 2;

    const int g_init = persistent_grid_for((const void*)k_pt_init, kBlock);
    const int g_ext_p = persistent_grid_for((const void*)k_pt_extend<true>, kBlock);
    const int g_ext_e = persistent_grid_for((const void*)k_pt_extend<false>, kBlock);
    const int g_shade = persistent_grid_for((const void*)k_pt_shade, kBlock);
    const int g_wshade = persistent_grid_for((const void*)k_wh_shade, kBlock);
    const int g_sh_p = persistent_grid_for((const void*)k_pt_shadow<true>, kBlock);
    const int g_sh_e = persistent_grid_for((const void*)k_pt_shadow<false>, kBlock);
    const int g_ext_c = persistent_grid_for((const void*)k_pt_extend_count<false>, kBlock);
    const int g_sh_c = persistent_grid_for((const void*)k_pt_shadow_count<false>, kBlock);

    // initial fill: sub-pool j starts with samples [first, first + n0_j)
    unsigned long long first = 0;
    for (int j = 0; j < plan.k; j++) {
        wrt_wavefront* wf = sub[j].wf;
        const unsigned long long left = P.total_samples - first;
        const unsigned n0 = (unsigned)std::min<unsigned long long>((unsigned long long)plan.cap[j], left);
        WRT_CUDA(cudaMemsetAsync(wf->counters, 0, WF_COUNTERS * sizeof(unsigned long long), st));
        if (whitted) k_wh_init<<<g_init, kBlock, 0, st>>>(P, dc, wf->pool, wf->queue[0], n0, first, wf->counters);
        else k_pt_init<<<g_init, kBlock, 0, st>>>(P, dc, wf->pool, wf->queue[0], n0, first, wf->counters);
        WRT_CUDA(cudaGetLastError());
        sub[j].index = j; sub[j].iter = 0; sub[j].batches = 0; sub[j].polled = 0; sub[j].last_n = n0; sub[j].done = n0 == 0; sub[j].timed = 0;
        first += n0;
        sc->stats.kernel_launches += 1;
    }
    unsigned long long* next_sample = &sub[0].wf->counters[WF_NEXT_SAMPLE];
    WRT_CUDA(cudaMemcpyAsync(next_sample, &first, sizeof first, cudaMemcpyHostToDevice, st));
    // fork: the sub-pool streams start after everything queued on the caller's stream so far
    WRT_CUDA(cudaEventRecord(sc->ev_fork, st));
    for (int j = 0; j < plan.k; j++) WRT_CUDA(cudaStreamWaitEvent(sub[j].wf->stream, sc->ev_fork, 0));

    const int kMaxTimed = 2048;
    const bool shadow_on_own_stream = !(getenv("WRT_SHADOW_STREAM") && atoi(getenv("WRT_SHADOW_STREAM")) == 0);   // A/B knob
    const bool regenerates = P.total_samples > first;       // some slot will take a second camera sample
    // Without regeneration a path slot lives for at most max_depth + 1 vertices (+ the emitter hit that ends it): the number
    // of iterations is known up front and nothing has to be read back at all.
    const int fixed_iters = regenerates ? 0 : (whitted ? 0 : P.max_depth + 2);

    // one iteration of sub-pool s: extend -> shade -> shadow on the sub-pool's stream, all lengths read on the device
    auto enqueue_iteration = [&](SubState& s) -> int {
        wrt_wavefront* wf = s.wf;
        cudaStream_t q = wf->stream;
        const int par = s.iter & 1, cur = s.iter & 1;           // queue[cur] is consumed, queue[cur ^ 1] produced
        const size_t cap = (size_t)wf->capacity;
        const bool time_it = s.timed < kMaxTimed;
        if (time_it && kEvPerIter * (s.timed + 1) > wf->n_ev) { int r = wavefront_events(wf, std::min(kEvPerIter * kMaxTimed, wf->n_ev * 2)); if (r) return r; }
        cudaEvent_t* ev = time_it ? &wf->ev[kEvPerIter * s.timed] : nullptr;
        cudaStream_t qs = shadow_on_own_stream ? wf->shadow_stream : q;
        const ShadowQueue& sq = par ? wf->shadow2 : wf->shadow;
        // bank `par` and shadow queue `par` were last used by iteration iter - 2, whose shadow kernel may still be running
        if (s.iter >= 2) WRT_CUDA(cudaStreamWaitEvent(q, wf->shadowed_ev[par], 0));
        WRT_CUDA(cudaMemsetAsync(wf->counters + par * WF_BANK, 0, WF_BANK * sizeof(unsigned long long), q));
        if (ev) cudaEventRecord(ev[0], q);
        if (counting && count_pruned) k_pt_extend_count<true><<<g_ext_c, kBlock, 0, q>>>(sc->view, wf->pool, wf->queue[cur], wf->counters, par, cap);
        else if (counting) k_pt_extend_count<false><<<g_ext_c, kBlock, 0, q>>>(sc->view, wf->pool, wf->queue[cur], wf->counters, par, cap);
        else if (pruned) k_pt_extend<true><<<g_ext_p, kBlock, 0, q>>>(sc->view, wf->pool, wf->queue[cur], wf->counters, par, (float4*)wf->trav_scratch, cap);
        else k_pt_extend<false><<<g_ext_e, kBlock, 0, q>>>(sc->view, wf->pool, wf->queue[cur], wf->counters, par, (float4*)wf->trav_scratch, cap);
        if (ev) cudaEventRecord(ev[1], q);
        if (whitted) k_wh_shade<<<g_wshade, kBlock, 0, q>>>(sc->view, P, dc, wf->pool, wf->queue[cur], wf->queue[cur ^ 1], sq,
                                                            d_film, wf->counters, par, next_sample, cap, wh_pend[s.index], wh_levels);
        else k_pt_shade<<<g_shade, kBlock, 0, q>>>(sc->view, P, dc, wf->pool, wf->queue[cur], wf->queue[cur ^ 1], sq,
                                                   d_film, wf->counters, par, next_sample, cap);
        if (ev) cudaEventRecord(ev[2], q);
        WRT_CUDA(cudaEventRecord(wf->shaded_ev[par], q));
        // the shadow kernel: its own stream, overlapping the next iteration's extend + shade
        WRT_CUDA(cudaStreamWaitEvent(qs, wf->shaded_ev[par], 0));
        if (ev) cudaEventRecord(ev[3], qs);
        if (counting && count_pruned) k_pt_shadow_count<true><<<g_sh_c, kBlock, 0, qs>>>(sc->view, sq, d_film, P.film_scale, wf->counters, par);
        else if (counting) k_pt_shadow_count<false><<<g_sh_c, kBlock, 0, qs>>>(sc->view, sq, d_film, P.film_scale, wf->counters, par);
        else if (pruned) k_pt_shadow<true><<<g_sh_p, kBlock, 0, qs>>>(sc->view, sq, d_film, P.film_scale, wf->counters, par, (float4*)wf->trav_scratch2);
        else k_pt_shadow<false><<<g_sh_e, kBlock, 0, qs>>>(sc->view, sq, d_film, P.film_scale, wf->counters, par, (float4*)wf->trav_scratch2);
        if (ev) { cudaEventRecord(ev[4], qs); s.timed++; }
        WRT_CUDA(cudaEventRecord(wf->shadowed_ev[par], qs));
        WRT_CUDA(cudaGetLastError());
        sc->stats.kernel_launches += 3;
        s.iter++;
        return WRT_OK;
    };
    // a batch of iterations, then the bank the last one wrote -> pinned slot (batch & 1), marked by an event
    auto enqueue_batch = [&](SubState& s) -> int {
        // large queues: one iteration per batch (an iteration takes milliseconds, nothing to gain); small ones: four,
        // so that short frames (torus.scene as shipped: 512 x 512 x 1 spp) never wait for the host
        const int n_it = s.last_n > (1u << 20) ? 1 : 4;
        for (int b = 0; b < n_it; b++) { int r = enqueue_iteration(s); if (r) return r; }
        wrt_wavefront* wf = s.wf;
        const int slot = s.batches & 1;
        const int last_par = (s.iter - 1) & 1;
        WRT_CUDA(cudaMemcpyAsync(wf->h_counters + slot * WF_BANK, wf->counters + last_par * WF_BANK, WF_BANK * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, wf->stream));
        WRT_CUDA(cudaEventRecord(wf->poll_ev[slot], wf->stream));
        s.batches++;
        return WRT_OK;
    };

    if (fixed_iters > 0) {
        for (int j = 0; j < plan.k; j++)
            if (!sub[j].done) for (int b = 0; b < fixed_iters; b++) { rc = enqueue_iteration(sub[j]); if (rc) return rc; }
    } else {
        // two batches in flight per sub-pool; whenever the older one's result is in, decide whether to enqueue another
        for (int j = 0; j < plan.k; j++) if (!sub[j].done) { rc = enqueue_batch(sub[j]); if (rc) return rc; }
        for (int j = 0; j < plan.k; j++) if (!sub[j].done) { rc = enqueue_batch(sub[j]); if (rc) return rc; }
        unsigned long long guard = 0;
        for (;;) {
            bool any = false;
            for (int j = 0; j < plan.k; j++) {
                SubState& s = sub[j];
                if (s.done) continue;
                any = true;
                const int slot = s.polled & 1;
                WRT_CUDA(cudaEventSynchronize(s.wf->poll_ev[slot]));
                s.last_n = wf_queue_n(s.wf->h_counters + slot * WF_BANK);
                s.polled++;
                if (s.last_n == 0) { s.done = true; continue; }      // (the younger batch in flight finds empty queues and returns at once)
                rc = enqueue_batch(s); if (rc) return rc;
            }
            if (!any) break;
            if (++guard > (1ull << 32)) { set_error("wrt_render_pt: runaway iteration count"); return WRT_ERR_CUDA; }
        }
    }
    // join: the caller's stream continues after every sub-pool stream (and its shadow stream)
    for (int j = 0; j < plan.k; j++) {
        for (int par = 0; par < 2; par++) if (sub[j].iter > par) WRT_CUDA(cudaStreamWaitEvent(sub[j].wf->stream, sub[j].wf->shadowed_ev[par], 0));
        WRT_CUDA(cudaEventRecord(sub[j].wf->join_ev, sub[j].wf->stream));
        WRT_CUDA(cudaStreamWaitEvent(st, sub[j].wf->join_ev, 0));
    }
    // ray counts of the render: accumulated on the device, read once
    sc->stats.extend_launches = 0; sc->stats.extend_rays = 0;
    for (int j = 0; j < plan.k; j++) {
        unsigned long long h[3];
        WRT_CUDA(cudaMemcpyAsync(h, &sub[j].wf->counters[WF_TOTAL_CLOSEST], sizeof h, cudaMemcpyDeviceToHost, st));
        WRT_CUDA(cudaStreamSynchronize(st));
        sc->stats.closest_rays += h[0]; sc->stats.shadow_rays += h[1];
        sc->stats.extend_rays += h[0]; sc->stats.extend_launches += h[2];
    }
    // stage times: sum of per-launch CUDA-event durations over all sub-pools (launches of different
    // sub-pools overlap, so the sums can exceed the wall time of the render)
    double ext = 0, shd = 0, shw = 0;
    for (int j = 0; j < plan.k; j++) {
        wavefront_sum_stage_times(sc, sub[j].wf, sub[j].timed);
        ext += sc->stats.extend_ms; shd += sc->stats.shade_ms; shw += sc->stats.shadow_ms;
    }
    sc->stats.extend_ms = ext; sc->stats.shade_ms = shd; sc->stats.shadow_ms = shw;
    if (counting) {
        for (int j = 0; j < plan.k; j++) {
            unsigned long long h[4];
            WRT_CUDA(cudaMemcpy(h, &sub[j].wf->counters[WF_VISITS], sizeof h, cudaMemcpyDeviceToHost));
            sc->stats.inner_visits += h[0]; sc->stats.leaf_visits += h[1]; sc->stats.tri_tests += h[2]; sc->stats.sphere_tests += h[3];
        }
    }
    sc->stats.samples += P.total_samples;
    return WRT_OK;
}

}  // namespace wrt

using namespace wrt;

extern "C" {

int wrt_render_pt_dev(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* d_film, void* stream)
{
    if (!sc || !d_film) { set_error("wrt_render_pt_dev: null argument"); return WRT_ERR_INVALID; }
    WRT_CUDA(cudaSetDevice(sc->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : sc->stream;
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    int rc = render_pt_device(sc, cam, p, d_film, st, false);
    if (rc) return rc;
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1)); sc->stats.last_render_ms = ms;
    return WRT_OK;
}

int wrt_render_whitted_dev(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* d_film, void* stream)
{
    if (!sc || !d_film) { set_error("wrt_render_whitted_dev: null argument"); return WRT_ERR_INVALID; }
    WRT_CUDA(cudaSetDevice(sc->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : sc->stream;
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    int rc = render_pt_device(sc, cam, p, d_film, st, true);
    if (rc) return rc;
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1)); sc->stats.last_render_ms = ms;
    return WRT_OK;
}

int wrt_film_resolve_dev(const float* d_film, int32_t width, int32_t height, float scale, float gamma, uint8_t* d_rgb, void* stream)
{
    if (!d_film || !d_rgb || width <= 0 || height <= 0 || !(gamma > 0.f)) { set_error("wrt_film_resolve_dev: bad argument"); return WRT_ERR_INVALID; }
    const size_t n = (size_t)width * height * 3;
    const int block = 256;
    const int grid = (int)std::min<size_t>((n + block - 1) / block, 148 * 8);
    k_film_resolve<<<grid, block, 0, (cudaStream_t)stream>>>(d_film, n, scale, 1.f / gamma, d_rgb);
    WRT_CUDA(cudaGetLastError());
    return WRT_OK;
}

static int render_pt_host(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* film, bool whitted);

int wrt_render_pt(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* film)
{
    return render_pt_host(sc, cam, p, film, false);
}

int wrt_render_whitted(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* film)
{
    return render_pt_host(sc, cam, p, film, true);
}

static int render_pt_host(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* film, bool whitted)
{
    if (!sc || !film || !p) { set_error("wrt_render_pt: null argument"); return WRT_ERR_INVALID; }
    WRT_CUDA(cudaSetDevice(sc->device));
    PtParams P;
    int rc = pt_fill_params(p, P);
    if (rc) return rc;
    if (sc->n_replicas > 0 && !sc->d_rng_tape && P.local_spp > 1) {
        // several devices (wrt_init): sample i of this call (k = sample_first + i * stride) goes to device i mod N
        const size_t floats = (size_t)p->width * p->height * 3;
        const int stride0 = P.sample_stride;
        auto fn = [&](wrt_scene* rs, int g, int n, float* d_film) -> int {
            wrt_pt_params q = *p;
            q.sample_first = P.sample_first + g * stride0;
            q.sample_stride = stride0 * n;
            q.film_scale = P.film_scale;                        // 1/spp of the WHOLE render, fused into every film_add
            return render_pt_device(rs, cam, &q, d_film, rs->stream, whitted);
        };
        float* film0 = nullptr;
        rc = multi_render(sc, floats, P.local_spp, fn, &film0);
        if (rc) return rc;
        WRT_CUDA(cudaMemcpyAsync(film, film0, floats * sizeof(float), cudaMemcpyDeviceToHost, sc->stream));
        WRT_CUDA(cudaStreamSynchronize(sc->stream));
        return WRT_OK;
    }
    const size_t floats = (size_t)p->width * p->height * 3;
    float* d_film = nullptr;
    rc = wavefront_film(sc, floats, &d_film);
    if (rc) return rc;
    cudaStream_t st = sc->stream;
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    WRT_CUDA(cudaMemsetAsync(d_film, 0, floats * sizeof(float), st));
    rc = render_pt_device(sc, cam, p, d_film, st, whitted);
    if (rc) return rc;
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaMemcpyAsync(film, d_film, floats * sizeof(float), cudaMemcpyDeviceToHost, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1)); sc->stats.last_render_ms = ms;
    return WRT_OK;
}

}  // extern "C"
