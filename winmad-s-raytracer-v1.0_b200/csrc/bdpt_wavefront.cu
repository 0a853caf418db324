// Wavefront bidirectional path tracer: wrt_render_bdpt, a drop-in for BidirPathTracing::render
// (R/src/surfaceIntegrator/bidirPathTracing.cpp:23-265).  Per iteration:
//   phase A  W*H light sub-paths: extend -> light_shade (store <= maxPathLength-1 vertices per path,
//            queue the connect-to-camera ray, scatter) -> connection kernel (occlusion + film splat)
//   phase B  W*H camera sub-paths: extend -> camera_shade (emitter hit, direct-illumination entry,
//            one connection query per stored light vertex of THIS pixel's light path, scatter)
//            -> connection kernel -> direct-illumination kernel (shadow ray, then the BSDF-sampled ray)
// Light vertices live in HBM as 4 x float4 (64 B) per vertex, [vertex k][path] so a warp's accesses
// to vertex k of consecutive paths are coalesced.  Per-vertex logic is bdpt_logic.cuh.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include "bdpt_logic.cuh"
#include "wavefront_kernels.cuh"

namespace wrt {

void fill_camera(const wrt_camera* c, DevCamera& d);

struct BdptBuffers {
    unsigned n_paths; int maxv;
    float* dvc;              // [n_paths] dVC of the live path
    float4* verts;           // [maxv][n_paths][4]
    int* nverts;             // [n_paths]
    ShadowQueue conn; size_t conn_cap;
    float4* di; size_t di_cap;   // 6 float4 per entry
    float4* cverts;          // [5][n_paths] camera-vertex records of the current iteration (k_bdpt_connect)
    uint32_t* pair_rec;      // [conn_cap] camera-vertex record of every (camera vertex, light vertex) pair, pairs of a record adjacent
};

__device__ __forceinline__ void bdpt_store(const PathPool& pool, float* dvc, uint32_t slot, const RayIn& r, const BdptPath& st)
{
    float4* p = reinterpret_cast<float4*>(pool.ray + slot);
    p[0] = make_float4(r.ox, r.oy, r.oz, r.dx);
    p[1] = make_float4(r.dy, r.dz, r.tmin, r.tmax);
    pool.weight_pdf[slot] = make_float4(st.throughput.x, st.throughput.y, st.throughput.z, st.dVCM);
    pool.meta[slot] = make_uint4(st.index, st.rng.key, st.rng.ctr, (uint32_t)st.length | ((uint32_t)st.spec << 16));
    dvc[slot] = st.dVC;
}

__device__ __forceinline__ void bdpt_load(const PathPool& pool, const float* dvc, uint32_t slot, BdptPath& st)
{
    const float4 w = pool.weight_pdf[slot];
    const uint4 m = pool.meta[slot];
    st.throughput = v3(w.x, w.y, w.z); st.dVCM = w.w; st.dVC = dvc[slot];
    st.index = m.x; st.rng.key = m.y; st.rng.ctr = m.z;
    st.length = (int)(m.w & 0xffffu); st.spec = (int)(m.w >> 16);
}

__device__ __forceinline__ void conn_store(const ShadowQueue& q, size_t pos, const Connection& c)
{
    q.a[pos] = make_float4(c.q[0], c.q[1], c.q[2], c.c.x);
    q.b[pos] = make_float4(c.q[3], c.q[4], c.q[5], c.c.y);
    q.c[pos] = make_float4(c.q[6], c.q[7], c.q[8], c.c.z);
    q.pixel[pos] = c.pixel;
}

__global__ void __launch_bounds__(kBlock)
k_bdpt_light_init(DevSceneView sc, BdptParams P, PathPool pool, BdptBuffers B, uint32_t* queue, unsigned long long* ctr, int first_parity)
{
    // the bank the phase's first iteration reads (device-driven loop, wavefront.h)
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctr[(first_parity ^ 1) * WF_BANK + WF_NEXT_COUNT] = P.n_paths; ctr[(first_parity ^ 1) * WF_BANK + WF_GEN_COUNT] = 0; }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_paths; i += gridDim.x * blockDim.x) {
        RayIn r; BdptPath st;
        bdpt_light_generate(sc, P, i, r, st);
        bdpt_store(pool, B.dvc, i, r, st);
        B.nverts[i] = 0;
        queue[i] = i;
    }
}

__global__ void __launch_bounds__(kBlock)
k_bdpt_camera_init(BdptParams P, DevCamera cam, PathPool pool, BdptBuffers B, uint32_t* queue, unsigned long long* ctr, int first_parity)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctr[(first_parity ^ 1) * WF_BANK + WF_NEXT_COUNT] = P.n_paths; ctr[(first_parity ^ 1) * WF_BANK + WF_GEN_COUNT] = 0; }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_paths; i += gridDim.x * blockDim.x) {
        RayIn r; BdptPath st;
        bdpt_camera_generate(P, cam, i, r, st);
        bdpt_store(pool, B.dvc, i, r, st);
        queue[i] = i;
    }
}

__global__ void __launch_bounds__(kBlock)
k_bdpt_light_shade(DevSceneView sc, BdptParams P, DevCamera cam, PathPool pool, BdptBuffers B,
                   const uint32_t* __restrict__ queue_in, uint32_t* __restrict__ queue_out,
                   unsigned long long* ctr, int parity)
{
    const size_t n = wf_queue_n(wf_prev(ctr, parity));
    unsigned long long* counters = wf_cur(ctr, parity);
    size_t base;
    while (next_chunk(&counters[WF_WORK2], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        const bool valid = e < n;
        uint32_t slot = 0;
        RayIn r; BdptPath st; LightStepOut out;
        out.alive = false; out.store = false; out.connect = false;
        if (valid) {
            slot = queue_in[e];
            r = pool_load_ray(pool, slot);
            bdpt_load(pool, B.dvc, slot, st);
            bdpt_light_step(sc, P, cam, r, st, pool.hit_prim[slot], pool.hit_t[slot], out);
            if (out.store) {
                const int k = B.nverts[slot];
                if (k < B.maxv) {
                    float4* v = B.verts + ((size_t)k * B.n_paths + slot) * 4;
                    v[0] = make_float4(out.v.pos.x, out.v.pos.y, out.v.pos.z, out.v.dVCM);
                    v[1] = make_float4(out.v.throughput.x, out.v.throughput.y, out.v.throughput.z, out.v.dVC);
                    v[2] = make_float4(out.v.wi.x, out.v.wi.y, out.v.wi.z, __int_as_float(out.v.matid));
                    v[3] = make_float4(out.v.n.x, out.v.n.y, out.v.n.z, __int_as_float(out.v.length | (out.v.spec << 16)));
                    B.nverts[slot] = k + 1;
                }
            }
        }
        const unsigned long long cpos = warp_append(&counters[WF_SHADOW_COUNT], valid && out.connect);
        if (valid && out.connect) conn_store(B.conn, cpos, out.conn);
        const bool alive = valid && out.alive;
        const unsigned long long qpos = warp_append(&counters[WF_NEXT_COUNT], alive);
        if (alive) { bdpt_store(pool, B.dvc, slot, r, st); queue_out[qpos] = slot; }
    }
}

// Camera-path vertex, split in two kernels (round 2; the single kernel held the path state, its BSDF, the direct-illumination
// entry AND a loop over the stored light vertices live at once: 168 registers, 12 resident warps per SM, 13.5 of 32 lanes
// busy — profiles/r2_ncu_bdpt_camera_shade_before.md):
//   k_bdpt_camera_shade  one thread per path: hit, emission, direct-illumination entry, the record the connections need
//                        (position, throughput, dVCM / dVC, wi, normal, material, length — everything connectVertices reads of the
//                        camera state, bidirPathTracing.cpp:611-665; the BSDF is rebuilt from (wi, n, material) exactly as it is for
//                        the stored light vertices), then sampleScattering;
//   k_bdpt_connect       one thread per (camera vertex, light vertex) pair.
// The reference walks the light vertices of a pixel in order and `break`s at the first one with
// lv.length + 1 + camera.length > maxPathLength (:238-240); stored lengths increase strictly along a light path, so the
// per-pair test `> max_len -> nothing` is the same set of pairs.
__global__ void __launch_bounds__(kBlock)
k_bdpt_camera_shade(DevSceneView sc, BdptParams P, PathPool pool, BdptBuffers B, const uint32_t* __restrict__ queue_in,
                    uint32_t* __restrict__ queue_out, float* __restrict__ film, unsigned long long* ctr, int parity)
{
    const size_t n = wf_queue_n(wf_prev(ctr, parity));
    unsigned long long* counters = wf_cur(ctr, parity);
    size_t base;
    while (next_chunk(&counters[WF_WORK2], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        const bool valid = e < n;
        uint32_t slot = 0;
        RayIn r; BdptPath st; Bsdf bsdf; V3 hit = v3(0, 0, 0), nrm = v3(0, 0, 0), emit_c = v3(0, 0, 0);
        bool emit = false, has_di = false;
        int k = 0;
        {
            DiEntry di;
            if (valid) {
                slot = queue_in[e];
                r = pool_load_ray(pool, slot);
                bdpt_load(pool, B.dvc, slot, st);
                k = bdpt_camera_pre(sc, P, r, st, pool.hit_prim[slot], pool.hit_t[slot], hit, nrm, bsdf, emit, emit_c, has_di, di);
                if (emit) film_add(film, st.index, emit_c, P.film_scale);
            }
            const unsigned long long dpos = warp_append(&counters[WF_AUX_COUNT], valid && has_di);
            if (valid && has_di) {
                float4* d = B.di + 6 * dpos;
                d[0] = make_float4(di.q[0], di.q[1], di.q[2], di.cA.x);
                d[1] = make_float4(di.q[3], di.q[4], di.q[5], di.cA.y);
                d[2] = make_float4(di.q[6], di.q[7], di.q[8], di.cA.z);
                d[3] = make_float4(di.bo[0], di.bo[1], di.bo[2], di.cB.x);
                d[4] = make_float4(di.bd[0], di.bd[1], di.bd[2], di.cB.y);
                d[5] = make_float4(__int_as_float(di.has_B), __int_as_float(di.light_id), __uint_as_float(di.pixel), di.cB.z);
            }
        }
        // the record of this vertex for the connections to the stored vertices of this pixel's light path
        const int nv = (valid && k == 1) ? B.nverts[slot] : 0;
        const unsigned long long vpos = warp_append(&counters[WF_CV_COUNT], nv > 0);
        // ... and its nv pairs in the pair list (warp scan of nv, one atomic per warp): the connection kernel then has one VALID
        // pair per lane (a [light vertex][record] grid left 12 of 32 lanes busy: most light paths are shorter than the longest)
        unsigned incl = (unsigned)nv;
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long pbase = 0;
        if (total) {
            if ((threadIdx.x & 31) == 31) pbase = atomicAdd(&counters[WF_PAIR_COUNT], (unsigned long long)total);
            pbase = __shfl_sync(0xffffffffu, pbase, 31) + (incl - (unsigned)nv);
        }
        if (nv > 0) {
            for (int q = 0; q < nv; q++) B.pair_rec[pbase + q] = (uint32_t)vpos;
            const size_t np = B.n_paths;
            B.cverts[0 * np + vpos] = make_float4(hit.x, hit.y, hit.z, st.dVCM);
            B.cverts[1 * np + vpos] = make_float4(st.throughput.x, st.throughput.y, st.throughput.z, st.dVC);
            B.cverts[2 * np + vpos] = make_float4(-r.dx, -r.dy, -r.dz, __int_as_float(bsdf.mat_id));
            B.cverts[3 * np + vpos] = make_float4(nrm.x, nrm.y, nrm.z, __int_as_float(st.length | (st.spec << 16)));
            B.cverts[4 * np + vpos] = make_float4(__uint_as_float(st.index), __uint_as_float(slot), __int_as_float(nv), __uint_as_float((uint32_t)pbase));
        }
        bool alive = false;
        if (valid && k != 0) {
            alive = bdpt_sample_scattering(sc, bsdf, hit, r, st);
            if (alive) st.length += 1;
        }
        const unsigned long long qpos = warp_append(&counters[WF_NEXT_COUNT], alive);
        if (alive) { bdpt_store(pool, B.dvc, slot, r, st); queue_out[qpos] = slot; }
    }
}

// Work items are uniform (one bsdf pair each), so they are dealt to the blocks statically — no work-fetch atomic — and the
// connection queue is appended to once per BLOCK pass (shared-memory scan of the four warp counts, one global atomic per 128
// items instead of one per 32): with ~10 connections per camera vertex the per-warp atomics on two addresses were what the
// kernel waited for.
__global__ void __launch_bounds__(kBlock)
k_bdpt_connect(DevSceneView sc, BdptParams P, BdptBuffers B, unsigned long long* ctr, int parity)
{
    unsigned long long* counters = wf_cur(ctr, parity);
    const size_t n = (size_t)counters[WF_PAIR_COUNT];      // work item i = pair i of the list k_bdpt_camera_shade wrote
    if (n == 0) return;
    const size_t np = B.n_paths;
    __shared__ unsigned s_cnt[kBlock / 32];
    __shared__ unsigned long long s_base;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t i0 = (size_t)blockIdx.x * kBlock; i0 < n; i0 += (size_t)gridDim.x * kBlock) {      // block-uniform trip count
        const size_t i = i0 + threadIdx.x;
        bool has = false; Connection c;
        if (i < n) {
            const size_t rec = B.pair_rec[i];
            const float4 c4 = B.cverts[4 * np + rec];
            const int v = (int)((uint32_t)i - __float_as_uint(c4.w));          // the record's pairs are adjacent: its light vertices 0 .. nv-1
            {
                const uint32_t slot = __float_as_uint(c4.y);
                const float4* q = B.verts + ((size_t)v * np + slot) * 4;
                const float4 q3 = q[3], c3 = B.cverts[3 * np + rec];
                const int ls = __float_as_int(q3.w), cs = __float_as_int(c3.w);
                const int lv_len = ls & 0xffff, cam_len = cs & 0xffff;
                if (lv_len + 1 + cam_len <= P.max_len && lv_len + 1 + cam_len >= P.min_len) {
                    const float4 q0 = q[0], q1 = q[1], q2 = q[2];
                    const float4 c0 = B.cverts[0 * np + rec], c1 = B.cverts[1 * np + rec], c2 = B.cverts[2 * np + rec];
                    LightVertex lv;
                    lv.pos = v3(q0.x, q0.y, q0.z); lv.dVCM = q0.w; lv.throughput = v3(q1.x, q1.y, q1.z); lv.dVC = q1.w;
                    lv.wi = v3(q2.x, q2.y, q2.z); lv.matid = __float_as_int(q2.w); lv.n = v3(q3.x, q3.y, q3.z);
                    lv.length = lv_len; lv.spec = ls >> 16;
                    BdptPath st;
                    st.throughput = v3(c1.x, c1.y, c1.z); st.dVCM = c0.w; st.dVC = c1.w;
                    st.length = cam_len; st.spec = cs >> 16; st.index = __float_as_uint(c4.x);
                    Bsdf bsdf;
                    bsdf_init(bsdf, v3(c2.x, c2.y, c2.z), v3(c3.x, c3.y, c3.z), __float_as_int(c2.w), sc);
                    has = bdpt_connect_vertices(sc, P, lv, bsdf, v3(c0.x, c0.y, c0.z), st, c);
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (lane == 0) s_cnt[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
            for (int w = 0; w < kBlock / 32; w++) { const unsigned t = s_cnt[w]; s_cnt[w] = tot; tot += t; }
            s_base = tot ? atomicAdd(&counters[WF_SHADOW_COUNT], (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (has) conn_store(B.conn, s_base + s_cnt[warp] + __popc(m & ((1u << lane) - 1u)), c);
        __syncthreads();      // s_cnt / s_base are rewritten by the next pass
    }
}

// getDirectIllumination's two scene queries: the light-sample shadow ray, then (only if that sample
// is visible, otherwise the outer weight is 0) the BSDF-sampled closest-hit ray.
template <bool PRUNED>
__global__ void __launch_bounds__(kBlock)
k_bdpt_di(DevSceneView sc, const float4* __restrict__ di, float* __restrict__ film, float scale, unsigned long long* ctr, int parity)
{
    unsigned long long* counters = wf_cur(ctr, parity);
    const size_t n = (size_t)counters[WF_AUX_COUNT];
    if (n == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctr[WF_TOTAL_SHADOW], (unsigned long long)n);
    size_t base;
    unsigned long long extra = 0;
    while (next_chunk(&counters[WF_WORK4], n, base)) {
        const size_t e = base + (threadIdx.x & 31);
        if (e >= n) continue;
        const float4* d = di + 6 * e;
        const float4 d0 = d[0], d1 = d[1], d2 = d[2];
        RayIn r;
        make_ray(d0.x, d0.y, d0.z, d1.x, d1.y, d1.z, r);
        if (!shadow_visible<PRUNED>(sc, r, d2.x, d2.y, d2.z)) continue;
        const float4 d5 = d[5];
        const uint32_t pixel = __float_as_uint(d5.z);
        film_add(film, pixel, v3(d0.w, d1.w, d2.w), scale);
        if (__float_as_int(d5.x)) {
            const float4 d3 = d[3], d4 = d[4];
            make_ray(d3.x, d3.y, d3.z, d4.x, d4.y, d4.z, r);
            float t;
            const int prim = kd_traverse<PRUNED, false>(sc, r, t, nullptr);
            extra++;
            if (prim >= 0) {
                const int m = __float_as_int(__ldg(&sc.prims[3 * (size_t)prim]).w);
                if (m < 0 && -m - 1 == __float_as_int(d5.y)) film_add(film, pixel, v3(d3.w, d4.w, d5.w), scale);
            }
        }
    }
    if (extra) atomicAdd(&ctr[WF_TOTAL_CLOSEST], extra);     // the BSDF-sampled closest-hit rays of getDirectIllumination
}

__global__ void k_transpose_film(const float* __restrict__ in, float* __restrict__ out, int n)
{
    const int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && j < n)
        for (int c = 0; c < 3; c++) out[3 * ((size_t)j * n + i) + c] = in[3 * ((size_t)i * n + j) + c];
}

void bdpt_destroy(wrt_wavefront* wf)
{
    BdptBuffers* B = (BdptBuffers*)wf->bdpt;
    if (!B) return;
    cudaFree(B->dvc); cudaFree(B->verts); cudaFree(B->nverts);
    cudaFree(B->conn.a); cudaFree(B->conn.b); cudaFree(B->conn.c); cudaFree(B->conn.pixel); cudaFree(B->di); cudaFree(B->cverts); cudaFree(B->pair_rec);
    delete B;
    wf->bdpt = nullptr;
}

static int bdpt_alloc(BdptBuffers* B, unsigned n_paths, int maxv)
{
    const size_t conn_cap = (size_t)n_paths * (size_t)(maxv + 1);
    WRT_CUDA(cudaMalloc((void**)&B->dvc, (size_t)n_paths * sizeof(float)));
    WRT_CUDA(cudaMalloc((void**)&B->verts, (size_t)n_paths * maxv * 4 * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&B->nverts, (size_t)n_paths * sizeof(int)));
    WRT_CUDA(cudaMalloc((void**)&B->conn.a, conn_cap * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&B->conn.b, conn_cap * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&B->conn.c, conn_cap * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&B->conn.pixel, conn_cap * sizeof(uint32_t)));
    WRT_CUDA(cudaMalloc((void**)&B->di, (size_t)n_paths * 6 * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&B->cverts, (size_t)n_paths * 5 * sizeof(float4)));
    WRT_CUDA(cudaMalloc((void**)&B->pair_rec, conn_cap * sizeof(uint32_t)));
    return WRT_OK;
}

// The buffers are published in wf->bdpt (with their sizes) only after every allocation has succeeded.
static int bdpt_buffers(wrt_wavefront* wf, unsigned n_paths, int maxv, BdptBuffers** out)
{
    BdptBuffers* B = (BdptBuffers*)wf->bdpt;
    if (B && B->n_paths >= n_paths && B->maxv >= maxv) {
        // the vertex array is indexed [k][n_paths]: keep the allocation's own stride
        *out = B; return WRT_OK;
    }
    bdpt_destroy(wf);
    B = new BdptBuffers();
    memset(B, 0, sizeof *B);
    const int rc = bdpt_alloc(B, n_paths, maxv);
    if (rc != WRT_OK) {
        wf->bdpt = B; bdpt_destroy(wf);      // frees the partial allocation, leaves wf->bdpt == nullptr
        cudaGetLastError();
        return rc;
    }
    B->n_paths = n_paths; B->maxv = maxv;
    B->conn_cap = (size_t)n_paths * (size_t)(maxv + 1);
    B->di_cap = n_paths;
    wf->bdpt = B;
    *out = B;
    return WRT_OK;
}

// Iterations are independent, so several run side by side in one wavefront: launches become `batch` times larger and
// `batch` times fewer (at 1440 x 1440 one iteration is only 2 M paths and ~60 sub-millisecond launches with a host round
// trip each).  Batch = as many iterations as fit 2^24 path slots (knob: WRT_BDPT_BATCH_PATHS).
static int bdpt_batch(const wrt_bdpt_params* p)
{
    const int stride = p->iter_stride > 0 ? p->iter_stride : 1;
    const int my_iters = std::max(1, (p->iterations - p->iter_first + stride - 1) / stride);
    unsigned long long slot_budget = 1ull << 24;
    if (const char* e = getenv("WRT_BDPT_BATCH_PATHS")) slot_budget = (unsigned long long)atoll(e);
    const unsigned long long npix = (unsigned long long)p->width * p->height;
    const int batch = (int)std::max<unsigned long long>(1ull, slot_budget / std::max<unsigned long long>(npix, 1ull));
    return std::min(batch, my_iters);
}

int render_bdpt_device(wrt_scene* sc, const wrt_camera* cam, const wrt_bdpt_params* p, float* d_film, cudaStream_t st)
{
    NvtxRange nvtx_range("wrt: BDPT wavefront");
    if (!cam || !p || p->width <= 0 || p->height <= 0 || p->iterations <= 0 || p->max_path_length < 1) {
        set_error("wrt_render_bdpt: bad parameters"); return WRT_ERR_INVALID;
    }
    if (p->width != p->height) {
        set_error("wrt_render_bdpt: the film must be square (the reference transposes it in place, bidirPathTracing.cpp:29-45)");
        return WRT_ERR_INVALID;
    }
    if (sc->view.n_lights <= 0) { set_error("wrt_render_bdpt: the scene has no light"); return WRT_ERR_INVALID; }
    if ((long long)p->width * p->height > (1ll << 26)) { set_error("wrt_render_bdpt: film too large"); return WRT_ERR_INVALID; }
    BdptParams P;
    P.width = p->width; P.height = p->height; P.min_len = p->min_path_length; P.max_len = p->max_path_length;
    P.control_len = p->control_length; P.seed = p->seed; P.iteration = 0;
    P.film_scale = p->film_scale != 0.f ? p->film_scale : 1.f / (float)p->iterations;
    P.n_pixels = (unsigned)(p->width * p->height);
    P.n_paths = P.n_pixels;
    P.light_path_num = (float)(p->width * p->height);
    { const char* e = getenv("WRT_BDPT_SKIP_GATED"); P.trace_gated = (e && atoi(e)) ? 0 : 1; }
    P.tape_stride = 0;
    struct TapeGuard {
        bool on = false;
        ~TapeGuard() { if (on) { const float* none = nullptr; cudaMemcpyToSymbol(c_rng_tape, &none, sizeof none); } }
    } tape_guard;
    if (sc->d_rng_tape) {      // diagnostics: replay the reference's random numbers (wrt_debug_set_rng_tape)
        const unsigned long long need = 2ull * P.n_pixels * (unsigned long long)p->iterations * sc->rng_tape_stride;
        if (need > sc->rng_tape_floats) { set_error("wrt_render_bdpt: the installed RNG tape is shorter than 2*width*height*iterations*stride"); return WRT_ERR_INVALID; }
        const float* tp = sc->d_rng_tape;
        WRT_CUDA(cudaMemcpyToSymbol(c_rng_tape, &tp, sizeof tp));
        tape_guard.on = true;
        P.tape_stride = sc->rng_tape_stride;
    }
    const int stride = p->iter_stride > 0 ? p->iter_stride : 1;
    if (p->iter_first < 0 || p->iter_first >= p->iterations) { set_error("wrt_render_bdpt: iter_first out of range"); return WRT_ERR_INVALID; }
    P.iter_stride = stride;
    const int maxv = std::max(P.max_len - 1, 1);
    if (maxv > 255) { set_error("wrt_render_bdpt: max_path_length > 256 is not supported (32-bit pair index of the connection kernel)"); return WRT_ERR_INVALID; }
    DevCamera dc; fill_camera(cam, dc);

    const int my_iters = (p->iterations - p->iter_first + stride - 1) / stride;
    const int batch = bdpt_batch(p);
    const unsigned max_slots = P.n_pixels * (unsigned)batch;

    wrt_wavefront* wf = nullptr;
    int rc = wavefront_get(sc, (int)std::max(max_slots, 1024u), &wf);
    if (rc) return rc;
    BdptBuffers* B = nullptr;
    rc = bdpt_buffers(wf, max_slots, maxv, &B);
    if (rc) return rc;
    BdptBuffers Bv = *B;
    Bv.n_paths = B->n_paths;      // stride of the vertex array
    const bool pruned = sc->traversal_mode == WRT_TRAVERSE_PRUNED;
    const bool counting = sc->counting != 0;
    const bool count_pruned = sc->counting == 2;

    const int g_li = persistent_grid_for((const void*)k_bdpt_light_init, kBlock);
    const int g_ci = persistent_grid_for((const void*)k_bdpt_camera_init, kBlock);
    const int g_ext_p = persistent_grid_for((const void*)k_pt_extend<true>, kBlock);
    const int g_ext_e = persistent_grid_for((const void*)k_pt_extend<false>, kBlock);
    const int g_ext_c = persistent_grid_for((const void*)k_pt_extend_count<false>, kBlock);
    const int g_ls = persistent_grid_for((const void*)k_bdpt_light_shade, kBlock);
    const int g_cs = persistent_grid_for((const void*)k_bdpt_camera_shade, kBlock);
    const int g_cn = persistent_grid_for((const void*)k_bdpt_connect, kBlock);
    const int g_sh_p = persistent_grid_for((const void*)k_pt_shadow<true>, kBlock);
    const int g_sh_e = persistent_grid_for((const void*)k_pt_shadow<false>, kBlock);
    const int g_sh_c = persistent_grid_for((const void*)k_pt_shadow_count<false>, kBlock);
    const int g_di_p = persistent_grid_for((const void*)k_bdpt_di<true>, kBlock);
    const int g_di_e = persistent_grid_for((const void*)k_bdpt_di<false>, kBlock);

    rc = wavefront_events(wf, kEvPerIter * 64);
    if (rc) return rc;
    const int kMaxTimed = 4096;
    int timed = 0;
    sc->stats.extend_launches = 0; sc->stats.extend_rays = 0;
    WRT_CUDA(cudaMemsetAsync(wf->counters, 0, WF_COUNTERS * sizeof(unsigned long long), st));
    // A sub-path has at most max_len vertices, so a phase is at most max_len + 1 iterations: the loop is enqueued whole,
    // every kernel reads its queue length from the bank its predecessor wrote, and nothing is read back until the end
    // (r1 synchronised and copied counters after each of ~2 x 11 iterations per batch).
    const int phase_iters = P.max_len + 1;
    int iter = 0;                                             // global iteration index: parity selects the counter bank

    for (int done_iters = 0; done_iters < my_iters; done_iters += batch) {
        const int nb = std::min(batch, my_iters - done_iters);
        P.iteration = p->iter_first + done_iters * stride;
        P.n_paths = P.n_pixels * (unsigned)nb;
        for (int phase = 0; phase < 2; phase++) {
            int cur = 0;
            if (phase == 0) k_bdpt_light_init<<<g_li, kBlock, 0, st>>>(sc->view, P, wf->pool, Bv, wf->queue[0], wf->counters, iter & 1);
            else k_bdpt_camera_init<<<g_ci, kBlock, 0, st>>>(P, dc, wf->pool, Bv, wf->queue[0], wf->counters, iter & 1);
            WRT_CUDA(cudaGetLastError());
            sc->stats.kernel_launches += 1;
            for (int it = 0; it < phase_iters; it++, iter++) {
                const int par = iter & 1;
                const bool time_it = timed < kMaxTimed;
                if (time_it && kEvPerIter * (timed + 1) > wf->n_ev) { rc = wavefront_events(wf, std::min(kEvPerIter * kMaxTimed, wf->n_ev * 2)); if (rc) return rc; }
                cudaEvent_t* ev = time_it ? &wf->ev[kEvPerIter * timed] : nullptr;
                WRT_CUDA(cudaMemsetAsync(wf->counters + par * WF_BANK, 0, WF_BANK * sizeof(unsigned long long), st));
                if (ev) cudaEventRecord(ev[0], st);
                if (counting && count_pruned) k_pt_extend_count<true><<<g_ext_c, kBlock, 0, st>>>(sc->view, wf->pool, wf->queue[cur], wf->counters, par, 0);
                else if (counting) k_pt_extend_count<false><<<g_ext_c, kBlock, 0, st>>>(sc->view, wf->pool, wf->queue[cur], wf->counters, par, 0);
                else if (pruned) k_pt_extend<true><<<g_ext_p, kBlock, 0, st>>>(sc->view, wf->pool, wf->queue[cur], wf->counters, par, (float4*)wf->trav_scratch, 0);
                else k_pt_extend<false><<<g_ext_e, kBlock, 0, st>>>(sc->view, wf->pool, wf->queue[cur], wf->counters, par, (float4*)wf->trav_scratch, 0);
                if (ev) cudaEventRecord(ev[1], st);
                if (phase == 0)
                    k_bdpt_light_shade<<<g_ls, kBlock, 0, st>>>(sc->view, P, dc, wf->pool, Bv, wf->queue[cur], wf->queue[cur ^ 1], wf->counters, par);
                else {
                    k_bdpt_camera_shade<<<g_cs, kBlock, 0, st>>>(sc->view, P, wf->pool, Bv, wf->queue[cur], wf->queue[cur ^ 1], d_film, wf->counters, par);
                    k_bdpt_connect<<<g_cn, kBlock, 0, st>>>(sc->view, P, Bv, wf->counters, par);
                }
                if (ev) { cudaEventRecord(ev[2], st); cudaEventRecord(ev[3], st); }
                if (counting && count_pruned) k_pt_shadow_count<true><<<g_sh_c, kBlock, 0, st>>>(sc->view, Bv.conn, d_film, P.film_scale, wf->counters, par);
                else if (counting) k_pt_shadow_count<false><<<g_sh_c, kBlock, 0, st>>>(sc->view, Bv.conn, d_film, P.film_scale, wf->counters, par);
                else if (pruned) k_pt_shadow<true><<<g_sh_p, kBlock, 0, st>>>(sc->view, Bv.conn, d_film, P.film_scale, wf->counters, par, (float4*)wf->trav_scratch);
                else k_pt_shadow<false><<<g_sh_e, kBlock, 0, st>>>(sc->view, Bv.conn, d_film, P.film_scale, wf->counters, par, (float4*)wf->trav_scratch);
                if (phase == 1) {
                    if (pruned) k_bdpt_di<true><<<g_di_p, kBlock, 0, st>>>(sc->view, Bv.di, d_film, P.film_scale, wf->counters, par);
                    else k_bdpt_di<false><<<g_di_e, kBlock, 0, st>>>(sc->view, Bv.di, d_film, P.film_scale, wf->counters, par);
                }
                if (ev) { cudaEventRecord(ev[4], st); timed++; }
                WRT_CUDA(cudaGetLastError());
                sc->stats.kernel_launches += 3 + (phase == 1 ? 2 : 0);
                cur ^= 1;
                if ((it & 15) == 15 && it + 1 < phase_iters) {      // very long paths allowed (max_len > 15): look once every 16 iterations
                    WRT_CUDA(cudaMemcpyAsync(wf->h_counters, wf->counters + par * WF_BANK, WF_BANK * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                    WRT_CUDA(cudaStreamSynchronize(st));
                    if (wf_queue_n(wf->h_counters) == 0) { iter++; break; }
                }
            }
        }
        sc->stats.samples += P.n_paths;
    }
    {   // ray counts of the render: accumulated on the device, read once
        unsigned long long h[3];
        WRT_CUDA(cudaMemcpyAsync(h, &wf->counters[WF_TOTAL_CLOSEST], sizeof h, cudaMemcpyDeviceToHost, st));
        WRT_CUDA(cudaStreamSynchronize(st));
        sc->stats.closest_rays += h[0]; sc->stats.shadow_rays += h[1];
        sc->stats.extend_rays += h[0]; sc->stats.extend_launches += h[2];
    }
    wavefront_sum_stage_times(sc, wf, timed);
    if (counting) {
        unsigned long long h[4];
        WRT_CUDA(cudaMemcpyAsync(h, &wf->counters[WF_VISITS], sizeof h, cudaMemcpyDeviceToHost, st));
        WRT_CUDA(cudaStreamSynchronize(st));
        sc->stats.inner_visits += h[0]; sc->stats.leaf_visits += h[1]; sc->stats.tri_tests += h[2]; sc->stats.sphere_tests += h[3];
    }
    return WRT_OK;
}

}  // namespace wrt

using namespace wrt;

extern "C" {

int wrt_render_bdpt_dev(wrt_scene* sc, const wrt_camera* cam, const wrt_bdpt_params* p, float* d_film, void* stream)
{
    if (!sc || !d_film || !p) { set_error("wrt_render_bdpt_dev: null argument"); return WRT_ERR_INVALID; }
    WRT_CUDA(cudaSetDevice(sc->device));
    if (p->transpose_output) { set_error("wrt_render_bdpt_dev: transpose_output is only offered by the host-buffer call"); return WRT_ERR_INVALID; }
    cudaStream_t st = stream ? (cudaStream_t)stream : sc->stream;
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    int rc = render_bdpt_device(sc, cam, p, d_film, st);
    if (rc) return rc;
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1)); sc->stats.last_render_ms = ms;
    return WRT_OK;
}

int wrt_render_bdpt(wrt_scene* sc, const wrt_camera* cam, const wrt_bdpt_params* p, float* film)
{
    if (!sc || !film || !p) { set_error("wrt_render_bdpt: null argument"); return WRT_ERR_INVALID; }
    WRT_CUDA(cudaSetDevice(sc->device));
    if (p->width <= 0 || p->height <= 0 || p->width != p->height) {
        set_error("wrt_render_bdpt: the film must be square and non-empty (bidirPathTracing.cpp:29-45)"); return WRT_ERR_INVALID;
    }
    const size_t floats = (size_t)p->width * p->height * 3;
    {
        const int stride0 = p->iter_stride > 0 ? p->iter_stride : 1;
        const int my_iters = p->iterations > p->iter_first ? (p->iterations - p->iter_first + stride0 - 1) / stride0 : 0;
        if (sc->n_replicas > 0 && !sc->d_rng_tape && my_iters > 1 && !p->transpose_output) {
            // several devices (wrt_init): iteration i of this call goes to device i mod N
            auto fn = [&](wrt_scene* rs, int g, int n, float* d_film) -> int {
                wrt_bdpt_params q = *p;
                q.iter_first = p->iter_first + g * stride0;
                q.iter_stride = stride0 * n;
                q.film_scale = p->film_scale != 0.f ? p->film_scale : 1.f / (float)p->iterations;
                return render_bdpt_device(rs, cam, &q, d_film, rs->stream);
            };
            float* film0 = nullptr;
            int rcm = multi_render(sc, floats, my_iters, fn, &film0);
            if (rcm) return rcm;
            WRT_CUDA(cudaMemcpyAsync(film, film0, floats * sizeof(float), cudaMemcpyDeviceToHost, sc->stream));
            WRT_CUDA(cudaStreamSynchronize(sc->stream));
            return WRT_OK;
        }
    }
    int rc;
    float* d_film = nullptr;
    rc = wavefront_film(sc, floats * 2, &d_film);
    if (rc) return rc;
    cudaStream_t st = sc->stream;
    WRT_CUDA(cudaEventRecord(sc->ev0, st));
    WRT_CUDA(cudaMemsetAsync(d_film, 0, floats * sizeof(float), st));
    rc = render_bdpt_device(sc, cam, p, d_film, st);
    if (rc) return rc;
    float* src = d_film;
    if (p->transpose_output) {
        dim3 b(16, 16), g((p->width + 15) / 16, (p->width + 15) / 16);
        k_transpose_film<<<g, b, 0, st>>>(d_film, d_film + floats, p->width);
        WRT_CUDA(cudaGetLastError());
        sc->stats.kernel_launches += 1;
        src = d_film + floats;
    }
    WRT_CUDA(cudaEventRecord(sc->ev1, st));
    WRT_CUDA(cudaMemcpyAsync(film, src, floats * sizeof(float), cudaMemcpyDeviceToHost, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f; WRT_CUDA(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1)); sc->stats.last_render_ms = ms;
    return WRT_OK;
}

}  // extern "C"
