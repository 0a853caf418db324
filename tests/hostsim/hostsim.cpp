// TEST-ONLY library ("hostsim"): the product's __host__ __device__ per-ray / per-path code
// (csrc/traverse.cuh, pt_logic.cuh, bdpt_logic.cuh) compiled for the CPU with
// g++ -ffp-contract=off and driven by plain loops.  It lets `pytest -m "not gpu"` check the control
// flow and arithmetic of the CUDA kernels against the oracle on a box with no GPU.  It is NOT part of
// libwrt_b200.so, is never imported by the product package, and is not a CPU fallback.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "scene_layout.h"
#include "pt_logic.cuh"
#include "bdpt_logic.cuh"
#include "whitted_logic.cuh"
#include "shading_kat.cuh"

using namespace wrt;

namespace wrt { void set_error(const std::string&) {} }

struct HsScene { SceneLayout L; };

static void cam_fill(const wrt_camera* c, DevCamera& d)
{
    for (int a = 0; a < 3; a++) { d.pos[a] = c->pos[a]; d.forward[a] = c->forward[a]; }
    d.image_plane_dist = c->image_plane_dist; d.x_res = c->x_res; d.y_res = c->y_res;
    memcpy(d.r2w, c->raster_to_world, sizeof d.r2w);
    memcpy(d.w2r, c->world_to_raster, sizeof d.w2r);
}


static inline int hs_traverse(const DevSceneView& sc, const RayIn& r, int pruned, float& t)
{
    return pruned ? kd_traverse<true, false>(sc, r, t, nullptr) : kd_traverse<false, false>(sc, r, t, nullptr);
}

static inline bool hs_visible(const DevSceneView& sc, const RayIn& r, int pruned, float px, float py, float pz)
{
    return pruned ? shadow_visible<true>(sc, r, px, py, pz) : shadow_visible<false>(sc, r, px, py, pz);
}

extern "C" {

int hs_scene_create(const wrt_scene_desc* d, void** out, char* err256)
{
    HsScene* h = new HsScene();
    std::string err;
    if (!build_layout(d, h->L, err)) { if (err256) snprintf(err256, 256, "%s", err.c_str()); delete h; return 1; }
    layout_point_view_at_host(h->L);
    *out = h;
    return 0;
}

void hs_scene_destroy(void* h) { delete (HsScene*)h; }

// RNG replay (shading.cuh): the host build reads the tape through h_rng_tape; stride 0 / null removes it.
static uint32_t g_tape_stride = 0;
void hs_set_rng_tape(const float* tape, unsigned stride) { h_rng_tape = tape; g_tape_stride = tape ? stride : 0; }

// The shading known-answer entry (csrc/shading_kat.cuh) on the CPU build of the same code.
void hs_debug_shading(void* hv, const wrt_camera* cam, int what, int iparam, const float* in, size_t n, float* out)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    DevCamera dc; memset(&dc, 0, sizeof dc);
    if (cam) cam_fill(cam, dc);
    const int is = shading_kat_in_stride(what), os = shading_kat_out_stride(what);
    for (size_t e = 0; e < n; e++) shading_kat(sc, dc, what, iparam, in + (size_t)is * e, out + (size_t)os * e);
}

int hs_num_nodes(void* h) { return ((HsScene*)h)->L.n_nodes; }
long long hs_num_recs(void* h) { return ((HsScene*)h)->L.n_recs; }

// FNV-1a over the bytes of the node, leaf-record and primitive arrays of the layout (serial = parallel build_layout test)
unsigned long long hs_layout_digest(void* hv)
{
    const SceneLayout& L = ((HsScene*)hv)->L;
    unsigned long long h = 1469598103934665603ull;
    auto eat = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
    eat(L.nodes.data(), L.nodes.size() * sizeof(float4));
    eat(L.recs.data(), L.recs.size() * sizeof(float4));
    eat(L.prims.data(), L.prims.size() * sizeof(float4));
    return h;
}

void hs_trace_closest(void* hv, const wrt_ray* rays, size_t n, int pruned, int32_t* prim, float* t)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    for (size_t i = 0; i < n; i++) {
        RayIn r = { rays[i].ox, rays[i].oy, rays[i].oz, rays[i].dx, rays[i].dy, rays[i].dz, rays[i].tmin, rays[i].tmax };
        float tt;
        prim[i] = hs_traverse(sc, r, pruned, tt);
        if (t) t[i] = tt;
    }
}

void hs_trace_closest_full(void* hv, const wrt_ray* rays, size_t n, int pruned, int32_t* prim, float* t,
                           float* p3, float* n3, int32_t* inside, int32_t* matid)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    for (size_t i = 0; i < n; i++) {
        RayIn r = { rays[i].ox, rays[i].oy, rays[i].oz, rays[i].dx, rays[i].dy, rays[i].dz, rays[i].tmin, rays[i].tmax };
        float tt;
        int id = hs_traverse(sc, r, pruned, tt);
        prim[i] = id; t[i] = tt;
        HitInfo h = { 0, 0, 0, 0, 0, 0, 0, 0 };
        if (id >= 0) fill_hit(sc, id, r, tt, h);
        p3[3 * i] = h.px; p3[3 * i + 1] = h.py; p3[3 * i + 2] = h.pz;
        n3[3 * i] = h.nx; n3[3 * i + 1] = h.ny; n3[3 * i + 2] = h.nz;
        inside[i] = h.inside; matid[i] = h.matid;
    }
}

void hs_trace_occluded(void* hv, const float* q9, size_t n, int pruned, uint8_t* occ)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    for (size_t i = 0; i < n; i++) {
        const float* q = q9 + 9 * i;
        RayIn r;
        make_ray(q[0], q[1], q[2], q[3], q[4], q[5], r);
        bool vis = hs_visible(sc, r, pruned, q[6], q[7], q[8]);
        occ[i] = vis ? 0 : 1;
    }
}

void hs_count_visits(void* hv, const wrt_ray* rays, size_t n, int pruned, unsigned long long* out4)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    for (size_t i = 0; i < n; i++) {
        RayIn r = { rays[i].ox, rays[i].oy, rays[i].oz, rays[i].dx, rays[i].dy, rays[i].dz, rays[i].tmin, rays[i].tmax };
        VisitCounters vc = { 0, 0, 0, 0 };
        float tt;
        if (pruned) kd_traverse<true, true>(sc, r, tt, &vc); else kd_traverse<false, true>(sc, r, tt, &vc);
        out4[0] += vc.inner; out4[1] += vc.leaf; out4[2] += vc.tri; out4[3] += vc.sph;
    }
}

void hs_visits_per_ray(void* hv, const wrt_ray* rays, size_t n, int pruned, unsigned* leaf, unsigned* tri)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    for (size_t i = 0; i < n; i++) {
        RayIn r = { rays[i].ox, rays[i].oy, rays[i].oz, rays[i].dx, rays[i].dy, rays[i].dz, rays[i].tmin, rays[i].tmax };
        VisitCounters vc = { 0, 0, 0, 0 };
        float tt;
        if (pruned) kd_traverse<true, true>(sc, r, tt, &vc); else kd_traverse<false, true>(sc, r, tt, &vc);
        leaf[i] = vc.leaf; tri[i] = vc.tri + vc.sph;
    }
}

// Sequential driver of whitted_shade (the per-node code of k_wh_shade) with an explicit list of parked children.
void hs_render_whitted(void* hv, const wrt_camera* cam, const wrt_pt_params* p, int pruned, float* film,
                       unsigned long long* rays_out)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    PtParams P;
    P.width = p->width; P.height = p->height; P.spp = p->spp; P.max_depth = p->max_depth; P.seed = p->seed;
    P.strata = (int)std::sqrt((double)p->spp); if (P.strata < 1) P.strata = 1;
    P.sample_first = p->sample_first; P.sample_stride = p->sample_stride > 0 ? p->sample_stride : 1;
    P.local_spp = (P.spp - P.sample_first + P.sample_stride - 1) / P.sample_stride;
    P.film_scale = p->film_scale != 0.f ? p->film_scale : 1.f / (float)p->spp;
    P.total_samples = (unsigned long long)P.width * P.height * (unsigned long long)P.local_spp;
    P.tape_stride = g_tape_stride;
    DevCamera dc; cam_fill(cam, dc);
    unsigned long long nrays = 0;
    struct Item { RayIn r; float w; int dep; };
    std::vector<Item> todo;
    for (unsigned long long s = 0; s < P.total_samples; s++) {
        RayIn r; PathData pd;
        pt_generate(P, dc, s, r, pd);
        float* px = film + 3 * (size_t)pd.pixel;
        todo.clear();
        todo.push_back(Item{ r, 1.f, 0 });
        while (!todo.empty()) {
            const Item it = todo.back(); todo.pop_back();
            float t;
            const int prim = hs_traverse(sc, it.r, pruned, t);
            nrays++;
            WhittedOut out;
            whitted_shade(sc, P.max_depth, it.r, it.dep, it.w, pd.rng, prim, t, out);
            if (out.emit) { px[0] += out.emit_c.x * P.film_scale; px[1] += out.emit_c.y * P.film_scale; px[2] += out.emit_c.z * P.film_scale; }
            if (out.shadow) {
                RayIn sr;
                make_ray(out.q[0], out.q[1], out.q[2], out.q[3], out.q[4], out.q[5], sr);
                nrays++;
                if (hs_visible(sc, sr, pruned, out.q[6], out.q[7], out.q[8])) {
                    px[0] += out.shadow_c.x * P.film_scale; px[1] += out.shadow_c.y * P.film_scale; px[2] += out.shadow_c.z * P.film_scale;
                }
            }
            if (out.trans.valid) todo.push_back(Item{ out.trans.ray, out.trans.w, it.dep + 1 });
            if (out.reflect.valid) todo.push_back(Item{ out.reflect.ray, out.reflect.w, it.dep + 1 });
        }
    }
    if (rays_out) *rays_out = nrays;
}

// Sequential driver of the same pt_generate / pt_shade the CUDA kernels call.
// rays_out (optional): closest + shadow rays traced.
void hs_render_pt(void* hv, const wrt_camera* cam, const wrt_pt_params* p, int pruned, float* film,
                  unsigned long long* rays_out)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    PtParams P;
    P.width = p->width; P.height = p->height; P.spp = p->spp; P.max_depth = p->max_depth; P.seed = p->seed;
    P.strata = (int)std::sqrt((double)p->spp); if (P.strata < 1) P.strata = 1;
    P.sample_first = p->sample_first; P.sample_stride = p->sample_stride > 0 ? p->sample_stride : 1;
    P.local_spp = (P.spp - P.sample_first + P.sample_stride - 1) / P.sample_stride;
    P.film_scale = p->film_scale != 0.f ? p->film_scale : 1.f / (float)p->spp;
    P.total_samples = (unsigned long long)P.width * P.height * (unsigned long long)P.local_spp;
    P.tape_stride = g_tape_stride;
    DevCamera dc; cam_fill(cam, dc);
    unsigned long long nrays = 0;
    for (unsigned long long s = 0; s < P.total_samples; s++) {
        RayIn r; PathData pd;
        pt_generate(P, dc, s, r, pd);
        for (;;) {
            float t;
            int prim = hs_traverse(sc, r, pruned, t);
            nrays++;
            ShadeOut out;
            pt_shade(sc, P, r, pd, prim, t, out);
            float* px = film + 3 * (size_t)pd.pixel;
            if (out.emit) { px[0] += out.emit_c.x * P.film_scale; px[1] += out.emit_c.y * P.film_scale; px[2] += out.emit_c.z * P.film_scale; }
            if (out.shadow) {
                RayIn sr;
                make_ray(out.q[0], out.q[1], out.q[2], out.q[3], out.q[4], out.q[5], sr);
                nrays++;
                bool vis = hs_visible(sc, sr, pruned, out.q[6], out.q[7], out.q[8]);
                if (vis) { px[0] += out.shadow_c.x * P.film_scale; px[1] += out.shadow_c.y * P.film_scale; px[2] += out.shadow_c.z * P.film_scale; }
            }
            if (!out.alive) break;
        }
    }
    if (rays_out) *rays_out = nrays;
}

// Sequential driver of the same bdpt_* functions the CUDA kernels call: BidirPathTracing::runIteration.
// film: raw accumulator film[a][b] (not transposed), scaled by film_scale (default 1/iterations).
void hs_render_bdpt(void* hv, const wrt_camera* cam, const wrt_bdpt_params* p, int pruned, float* film,
                    unsigned long long* rays_out)
{
    const DevSceneView& sc = ((HsScene*)hv)->L.view;
    DevCamera dc; cam_fill(cam, dc);
    BdptParams P;
    P.width = p->width; P.height = p->height; P.min_len = p->min_path_length; P.max_len = p->max_path_length;
    P.control_len = p->control_length; P.seed = p->seed;
    P.film_scale = p->film_scale != 0.f ? p->film_scale : 1.f / (float)p->iterations;
    P.n_paths = (unsigned)(p->width * p->height); P.light_path_num = (float)(p->width * p->height);
    P.n_pixels = P.n_paths; P.iter_stride = 1;      // one iteration at a time here; the kernels batch several
    P.trace_gated = 1;
    P.tape_stride = g_tape_stride;
    const int stride = p->iter_stride > 0 ? p->iter_stride : 1;
    const int maxv = P.max_len > 1 ? P.max_len - 1 : 1;
    unsigned long long nrays = 0;
    std::vector<LightVertex> verts((size_t)P.n_paths * maxv);
    std::vector<int> nverts(P.n_paths);
    auto trace = [&](const RayIn& r, float& t) {
        nrays++;
        return pruned ? kd_traverse<true, false>(sc, r, t, nullptr) : kd_traverse<false, false>(sc, r, t, nullptr);
    };
    auto visible = [&](const float* q) {
        RayIn r; make_ray(q[0], q[1], q[2], q[3], q[4], q[5], r);
        nrays++;
        return pruned ? shadow_visible<true>(sc, r, q[6], q[7], q[8]) : shadow_visible<false>(sc, r, q[6], q[7], q[8]);
    };
    auto add = [&](uint32_t pixel, V3 c) {
        if (pixel == 0xffffffffu) return;
        float* px = film + 3 * (size_t)pixel;
        px[0] += c.x * P.film_scale; px[1] += c.y * P.film_scale; px[2] += c.z * P.film_scale;
    };
    for (int it = p->iter_first; it < p->iterations; it += stride) {
        P.iteration = it;
        for (uint32_t i = 0; i < P.n_paths; i++) {              // light paths
            RayIn ray; BdptPath st;
            bdpt_light_generate(sc, P, i, ray, st);
            nverts[i] = 0;
            for (;;) {
                float t; int prim = trace(ray, t);
                LightStepOut out;
                bdpt_light_step(sc, P, dc, ray, st, prim, t, out);
                if (out.store && nverts[i] < maxv) verts[(size_t)i * maxv + nverts[i]++] = out.v;
                if (out.connect && visible(out.conn.q)) add(out.conn.pixel, out.conn.c);
                if (!out.alive) break;
            }
        }
        for (uint32_t i = 0; i < P.n_paths; i++) {              // camera paths
            RayIn ray; BdptPath st;
            bdpt_camera_generate(P, dc, i, ray, st);
            for (;;) {
                float t; int prim = trace(ray, t);
                V3 hit, emit_c; Bsdf bsdf; bool emit, has_di; DiEntry di;
                V3 nrm; int k = bdpt_camera_pre(sc, P, ray, st, prim, t, hit, nrm, bsdf, emit, emit_c, has_di, di);
                if (emit) add(st.index, emit_c);
                if (k == 0) break;
                if (has_di && visible(di.q)) {
                    add(di.pixel, di.cA);
                    if (di.has_B) {
                        RayIn br; make_ray(di.bo[0], di.bo[1], di.bo[2], di.bd[0], di.bd[1], di.bd[2], br);
                        float tb; int pb = trace(br, tb);
                        if (pb >= 0) {
                            int m = f2i(sc.prims[3 * (size_t)pb].w);
                            if (m < 0 && -m - 1 == di.light_id) add(di.pixel, di.cB);
                        }
                    }
                }
                if (k == 1) {
                    for (int v = 0; v < nverts[i]; v++) {
                        const LightVertex& lv = verts[(size_t)i * maxv + v];
                        if (lv.length + 1 + st.length < P.min_len) continue;
                        if (lv.length + 1 + st.length > P.max_len) break;
                        Connection c;
                        if (bdpt_connect_vertices(sc, P, lv, bsdf, hit, st, c) && visible(c.q)) add(c.pixel, c.c);
                    }
                }
                if (!bdpt_sample_scattering(sc, bsdf, hit, ray, st)) break;
                st.length += 1;
            }
        }
    }
    if (rays_out) *rays_out = nrays;
}

}  // extern "C"
