// Bidirectional path tracing, one call per path vertex: the restatement of
//   BidirPathTracing::runIteration       R/src/surfaceIntegrator/bidirPathTracing.cpp:53-265
//   generateLightSample                  :267-311      connectToCamera        :313-368
//   sampleScattering                     :370-416      generateCameraSample   :418-452
//   getLightRadiance                     :454-482      getDirectIllumination  :484-608
//   connectVertices                      :611-665
// split at the reference's calls into the scene (scene.intersect / scene.occluded) so traversal runs
// as its own kernels.  The shipped behaviour is reproduced, including what SURVEY.md Appendix C lists:
//   * only paths of total length == controlLength are accumulated (:107-108,188-189,207-208,252-255),
//     while the connection rays of every (camera vertex, light vertex) pair are still traced (:244-245);
//   * MIS-weighted connection / direct-light terms are multiplied again by 1/(length - specular) (:211-215,248-255);
//   * getDirectIllumination's inner `weight` shadows the outer one: the function returns res * weight
//     with the OUTER weight, which is 0 unless the light-sampling branch found an unoccluded sample (:489,555,607),
//     and returns res un-weighted when the BSDF-sampled direction sees light pdf ~ 0 (:561-562) — with
//     invArea < EPS (any emitter larger than 1000 area units, e.g. torus.scene) that is always;
//   * the film is indexed [(int)x][(int)y] with x = pathIndex / width (transposed; :263,418-427,110-118);
//   * emitter vertices on light paths are terminal (their BSDF is uninitialised in the reference, :78-86 of
//     bsdf.h; every use of it evaluates to black).
// control_length == 0 is OUR extension meaning "no gating" (all path lengths contribute).
#pragma once
#include "shading.cuh"

namespace wrt {

struct BdptParams {
    int width, height;
    int min_len, max_len, control_len;
    uint32_t seed;
    int iteration;            // global index of the first iteration of this batch: keys the RNG streams
    int iter_stride;          // path slot i belongs to iteration `iteration + (i / n_pixels) * iter_stride`, pixel i % n_pixels
    float film_scale;
    float light_path_num;     // lightPathNum = width * height
    unsigned n_pixels;        // width * height: paths per iteration
    unsigned n_paths;         // path slots of this batch = (iterations in the batch) * n_pixels
    int trace_gated;          // 1: also trace connection rays whose contribution is gated out (as the reference does)
    uint32_t tape_stride;     // > 0: RNG replay; light path p of iteration it reads tape[((it*n_pixels + p)*2 + 0)*stride ...], camera path + 1
};

struct BdptPath {             // BidirPathState without its BSDF / origin / dir (those live in the ray)
    V3 throughput;
    float dVCM, dVC;
    int length, spec;
    Rng rng;
    uint32_t index;           // light: pathIndex; camera: film index (or 0xffffffff when off the film)
};

struct LightVertex {          // what connectVertices needs of a stored light BidirPathState
    V3 pos, throughput, wi, n;
    float dVCM, dVC;
    int matid, length, spec;
};

struct Connection {           // Scene::occluded query + what to add to the film when it is not occluded
    float q[9];
    V3 c;
    uint32_t pixel;
};

struct DiEntry {              // getDirectIllumination split at its two scene queries
    float q[9];               // light-sample shadow query
    V3 cA;                    // contribution of the light-sample branch (outer weight folded in)
    float bo[3], bd[3];       // BSDF-sampled ray Ray(hit + d*EPS, d)  (origin, raw direction)
    V3 cB;                    // its contribution if it hits light `light_id` (outer weight folded in)
    int has_B, light_id;
    uint32_t pixel;
};

WRT_HD bool bdpt_gate(const BdptParams& P, int len) { return P.control_len == 0 || len == P.control_len; }
WRT_HD V3 cdiv(V3 c, float s) { return v3(c.x / s, c.y / s, c.z / s); }   // Color3 operator/ (no epsilon guard)

// generateLightSample + the first Ray(origin + dir*EPS, dir) of the light loop (:73-80)
WRT_HD void bdpt_light_generate(const DevSceneView& sc, const BdptParams& P, uint32_t index, RayIn& ray, BdptPath& st)
{
    // several iterations run side by side (the reference runs them one after the other; they are independent: camera
    // path p only reads light path p of its own iteration, bidirPathTracing.cpp:222-229)
    const uint32_t within = index % P.n_pixels;
    const unsigned long long iter = (unsigned long long)P.iteration + (unsigned long long)(index / P.n_pixels) * (unsigned)P.iter_stride;
    st.rng = P.tape_stride ? rng_make_tape((iter * P.n_pixels + within) * 2ull, P.tape_stride) : rng_make(P.seed, 1u, iter * P.n_pixels + within);
    const int nl = sc.n_lights;
    const float pick = 1.f / nl;
    int lid = (int)(rng_float(st.rng) * nl);
    if (lid >= nl) lid = nl - 1;
    V3 pos = v3(0, 0, 0), dir = v3(0, 0, 1), rad = v3(0, 0, 0);
    float epdf = 0.f, dpdf = 0.f, cosl = 0.f;
    for (int tries = 0; tries < 64; tries++) {     // reference: for(;;) until emissionPdf > 1e-7
        // light->emit(sphere, rng.randVector3(), rng.randVector3(), ...) (:282-285): the order of the two draws is the
        // compiler's choice; g++ on x86-64 evaluates arguments right to left, so the POSITION numbers come first
        // (checked by the RNG-tape tests, which replay the compiled reference's stream)
        const V3 pr = rng_vec3(st.rng), dr = rng_vec3(st.rng);
        rad = light_emit(sc.lights[lid], dr, pr, pos, dir, epdf, dpdf, cosl);
        if (epdf > 1e-7f) break;
    }
    epdf *= pick; dpdf *= pick;
    st.throughput = cdiv(rad, epdf);
    st.length = 1; st.spec = 0;
    st.dVCM = dpdf / epdf;
    st.dVC = 1.f / epdf;                           // area lights are not delta
    st.index = index;
    make_ray(pos.x + dir.x * WRT_EPS, pos.y + dir.y * WRT_EPS, pos.z + dir.z * WRT_EPS, dir.x, dir.y, dir.z, ray);
}

// generateCameraSample + the first Ray(origin + dir*EPS, dir) of the camera loop (:147-149)
WRT_HD void bdpt_camera_generate(const BdptParams& P, const DevCamera& cam, uint32_t index, RayIn& ray, BdptPath& st)
{
    const uint32_t within = index % P.n_pixels;
    const unsigned long long iter = (unsigned long long)P.iteration + (unsigned long long)(index / P.n_pixels) * (unsigned)P.iter_stride;
    st.rng = P.tape_stride ? rng_make_tape((iter * P.n_pixels + within) * 2ull + 1ull, P.tape_stride) : rng_make(P.seed, 2u, iter * P.n_pixels + within);
    const int y = (int)(within % (uint32_t)P.width), x = (int)(within / (uint32_t)P.width);
    const V3 j = rng_vec3(st.rng);
    const float sx = (float)x + j.x, sy = (float)y + j.y;
    RayIn cr;
    camera_ray(cam, sx, sy, cr);
    const float cos_cam = cam.forward[0] * cr.dx + cam.forward[1] * cr.dy + cam.forward[2] * cr.dz;
    const float ipd = cam.image_plane_dist / cos_cam;
    const float cam_pdf = (ipd * ipd) / cos_cam;
    st.throughput = v3(1.f, 1.f, 1.f);
    st.length = 1; st.spec = 0;
    st.dVCM = P.light_path_num / cam_pdf;
    st.dVC = 0.f;
    const int fi = (int)sx, fj = (int)sy;           // film->addColor((int)sample.x, (int)sample.y, ...) :263
    st.index = (fi >= 0 && fi < P.height && fj >= 0 && fj < P.width) ? (uint32_t)(fi * P.width + fj) : 0xffffffffu;
    make_ray(cr.ox + cr.dx * WRT_EPS, cr.oy + cr.dy * WRT_EPS, cr.oz + cr.dz * WRT_EPS, cr.dx, cr.dy, cr.dz, ray);
}

// sampleScattering (:370-416) + the next Ray(origin + dir*EPS, dir)
WRT_HD bool bdpt_sample_scattering(const DevSceneView& sc, const Bsdf& bsdf, V3 hit, RayIn& ray, BdptPath& st)
{
    float dpdf = 0.f, cos_wo = 0.f; int type = 0;
    V3 wo = v3(ray.dx, ray.dy, ray.dz);
    const V3 f = bsdf_sample(bsdf, sc, rng_vec3(st.rng), wo, dpdf, cos_wo, type);
    if (is_black(f)) return false;
    float rpdf = dpdf;
    if ((type & BSDF_SPECULAR) == 0) rpdf = bsdf_pdf(bsdf, sc, wo, true);
    const float cont = bsdf.continue_prob;
    if (rng_float(st.rng) > cont) return false;
    dpdf *= cont; rpdf *= cont;
    if (type & BSDF_SPECULAR) {
        st.spec += 1;
        st.dVCM = 0.f;
        st.dVC *= cos_wo;
    } else {
        st.dVC = (1.f / dpdf) * (st.dVCM + st.dVC * rpdf);
        st.dVCM = 1.f / dpdf;
    }
    st.throughput = mul(st.throughput, f) * (cos_wo / dpdf);
    make_ray(hit.x + wo.x * WRT_EPS, hit.y + wo.y * WRT_EPS, hit.z + wo.z * WRT_EPS, wo.x, wo.y, wo.z, ray);
    return true;
}

// connectToCamera (:313-368) without its scene.occluded call: fills the query and the weighted result.
WRT_HD bool bdpt_connect_to_camera(const DevSceneView& sc, const BdptParams& P, const DevCamera& cam, const BdptPath& st,
                                   V3 hit, const Bsdf& bsdf, Connection& c)
{
    const V3 cpos = ld3(cam.pos), fwd = ld3(cam.forward);
    V3 d = cpos - hit;
    if (dot(-d, fwd) <= 0.f) return false;
    const float dist2 = sqr_length(d);
    const float dist = sqrtf(dist2);
    d = div_eps(d, dist);
    float cos_to_cam = 0.f, dpdf, rpdf;
    const V3 f = bsdf_f(bsdf, sc, d, cos_to_cam, &dpdf, &rpdf);
    if (is_black(f)) return false;
    rpdf *= bsdf.continue_prob;
    const float cos_cam = dot(-d, fwd);
    const float ipd = cam.image_plane_dist / cos_cam;
    const float i2sa = (ipd * ipd) / cos_cam;
    const float i2surf = i2sa * fabsf(cos_to_cam) / dist2;
    const float s2i = 1.f / i2surf;
    const V3 res = cdiv(mul(st.throughput, f), P.light_path_num * s2i);
    if (is_black(res)) return false;
    const float w_light = (i2surf / P.light_path_num) * (st.dVCM + rpdf * st.dVC);
    const float w = 1.f / (w_light + 1.f);
    c.q[0] = hit.x; c.q[1] = hit.y; c.q[2] = hit.z; c.q[3] = d.x; c.q[4] = d.y; c.q[5] = d.z;
    c.q[6] = cpos.x; c.q[7] = cpos.y; c.q[8] = cpos.z;
    c.c = res * w;
    return true;
}

struct LightStepOut {
    bool alive, store, connect;
    LightVertex v;
    Connection conn;
};

// One vertex of the light-path loop (:73-128), given the closest hit of `ray`.
WRT_HD void bdpt_light_step(const DevSceneView& sc, const BdptParams& P, const DevCamera& cam, RayIn& ray, BdptPath& st,
                            int prim, float t, LightStepOut& out)
{
    out.alive = false; out.store = false; out.connect = false;
    if (prim < 0) return;
    HitInfo h;
    fill_hit(sc, prim, ray, t, h);
    const V3 hit = v3(h.px, h.py, h.pz);
    const V3 rdir = v3(ray.dx, ray.dy, ray.dz);
    Bsdf bsdf;
    bsdf_init(bsdf, -rdir, v3(h.nx, h.ny, h.nz), h.matid, sc);
    if (bsdf.mat_id == 0) return;
    st.dVCM *= t * t;                                 // isFiniteLight is true for area lights (:90-93)
    const float acw = fabsf(bsdf.wi_local.z);
    st.dVCM /= acw; st.dVC /= acw;
    if (bsdf.mat_id < 0) return;                      // emitter vertex: terminal
    if (!bsdf.is_delta) {
        out.store = true;
        out.v.pos = hit; out.v.throughput = st.throughput; out.v.wi = -rdir; out.v.n = v3(h.nx, h.ny, h.nz);
        out.v.dVCM = st.dVCM; out.v.dVC = st.dVC; out.v.matid = h.matid; out.v.length = st.length; out.v.spec = st.spec;
        if (st.length + 1 >= P.min_len && bdpt_gate(P, st.length + 1)) {            // connect to camera :105-121
            const V3 ip = t_point(cam.w2r, hit);
            if (camera_check_raster(cam, ip.x, ip.y) && bdpt_connect_to_camera(sc, P, cam, st, hit, bsdf, out.conn)) {
                const int fi = (int)ip.x, fj = (int)ip.y;
                if (fi >= 0 && fi < P.height && fj >= 0 && fj < P.width) {          // ImageFilm::addColor bounds
                    out.conn.pixel = (uint32_t)(fi * P.width + fj);
                    out.connect = true;
                }
            }
        }
    }
    if (st.length + 2 > P.max_len) return;
    if (!bdpt_sample_scattering(sc, bsdf, hit, ray, st)) return;
    st.length += 1;
    out.alive = true;
}

// getLightRadiance (:454-482)
WRT_HD V3 bdpt_light_radiance(const DevSceneView& sc, const DevLight& L, const BdptPath& st, V3 ray_dir)
{
    const float pick = 1.f / sc.n_lights;
    float dpa, epdf;
    const V3 rad = light_radiance(L, ray_dir, &dpa, &epdf);
    if (is_black(rad)) return v3(0, 0, 0);
    if (st.length == 1) return rad;
    dpa *= pick; epdf *= pick;
    const float w_cam = dpa * st.dVCM + epdf * st.dVC;
    return rad * (1.f / (1.f + w_cam));
}

// getDirectIllumination (:484-608) up to its scene queries.  `scale` = throughput * 1/(len+1-spec).
// Returns false when nothing can reach the film (no query is queued).
WRT_HD bool bdpt_direct_illumination(const DevSceneView& sc, const Bsdf& bsdf, V3 hit, BdptPath& st, V3 scale, DiEntry& e)
{
    const int nl = sc.n_lights;
    const float pick = 1.f / nl;
    int lid = (int)(rng_float(st.rng) * nl);
    if (lid >= nl) lid = nl - 1;
    const DevLight& L = sc.lights[lid];
    V3 dl; float dist, dpdf, epdf, cosl;
    const V3 illu = light_illuminance(L, hit, rng_vec3(st.rng), dl, dist, dpdf, &epdf, &cosl);
    bool has_A = false;
    V3 colorA = v3(0, 0, 0);
    float w_outer = 0.f;
    if (!is_black(illu) && dpdf > 0.f) {
        float cos_l = 0.f, bd, br;
        const V3 f = bsdf_f(bsdf, sc, dl, cos_l, &bd, &br);
        if (!is_black(f)) {
            const float cont = bsdf.continue_prob;
            bd *= cont; br *= cont;
            const V3 tmp = cdiv(mul(illu, f) * cos_l, dpdf * pick);
            if (!is_black(tmp)) {
                const float w_light = bd / (dpdf * pick);
                const float w_cam = (epdf * cos_l / (dpdf * cosl)) * (st.dVCM + br * st.dVC);
                w_outer = 1.f / (w_light + 1.f + w_cam);
                const float wa = dpdf / (dpdf + bd);
                colorA = tmp * wa;
                has_A = true;
                const V3 p2 = hit + dl * dist;
                e.q[0] = hit.x; e.q[1] = hit.y; e.q[2] = hit.z; e.q[3] = dl.x; e.q[4] = dl.y; e.q[5] = dl.z;
                e.q[6] = p2.x; e.q[7] = p2.y; e.q[8] = p2.z;
            }
        }
    }
    // BSDF-sampled branch (:537-605); the RNG draw happens whether or not branch A succeeded
    bool quirk_unweighted = false, has_B = false;
    V3 colorB = v3(0, 0, 0), db = dl;
    {
        float pdf_b = 0.f, cos_s = 0.f; int type = 0;
        const V3 fb = bsdf_sample(bsdf, sc, rng_vec3(st.rng), db, pdf_b, cos_s, type);
        if (!is_black(fb) && pdf_b > 0.f) {
            float wi = 1.f;
            V3 illu_b = illu;                              // specular lobes reuse `illu` of branch A
            if ((type & BSDF_SPECULAR) == 0) {
                float light_pdf, e2;
                illu_b = light_radiance(L, db, &light_pdf, &e2);
                if (cmpf(light_pdf) == 0) quirk_unweighted = true;   // `return res;` — un-weighted (:561-562)
                else wi = pdf_b / (pdf_b + light_pdf);
            }
            if (!quirk_unweighted && !is_black(illu_b)) {
                colorB = cdiv(mul(illu_b, fb) * cos_s, pdf_b) * wi;
                has_B = true;
            }
        }
    }
    if (!has_A) return false;          // outer weight stays 0 (or res is 0 in the un-weighted return)
    const float wo = quirk_unweighted ? 1.f : w_outer;
    e.cA = mul(scale, colorA) * wo;
    e.has_B = (!quirk_unweighted && has_B) ? 1 : 0;
    e.cB = mul(scale, colorB) * wo;
    e.bo[0] = hit.x + db.x * WRT_EPS; e.bo[1] = hit.y + db.y * WRT_EPS; e.bo[2] = hit.z + db.z * WRT_EPS;
    e.bd[0] = db.x; e.bd[1] = db.y; e.bd[2] = db.z;
    e.light_id = lid;
    e.pixel = st.index;
    return true;
}

// Head of one camera-path vertex (:150-217).  Returns 0 = path ends, 1 = continue and connect to the
// light vertices (non-delta BSDF), 2 = continue without connections (delta BSDF).
WRT_HD int bdpt_camera_pre(const DevSceneView& sc, const BdptParams& P, const RayIn& ray, BdptPath& st, int prim, float t,
                           V3& hit, V3& nrm, Bsdf& bsdf, bool& emit, V3& emit_c, bool& has_di, DiEntry& di)
{
    emit = false; has_di = false;
    if (prim < 0) return 0;
    HitInfo h;
    fill_hit(sc, prim, ray, t, h);
    hit = v3(h.px, h.py, h.pz);
    nrm = v3(h.nx, h.ny, h.nz);      // the normal the BSDF was built from (the connection kernel rebuilds the BSDF from it)
    const V3 rdir = v3(ray.dx, ray.dy, ray.dz);
    bsdf_init(bsdf, -rdir, v3(h.nx, h.ny, h.nz), h.matid, sc);
    if (bsdf.mat_id == 0) return 0;
    st.dVCM *= t * t;
    const float acw = fabsf(bsdf.wi_local.z);
    st.dVCM /= acw; st.dVC /= acw;
    if (h.matid < 0) {                                                   // emitter hit :184-199
        const int li = -h.matid - 1;
        if (li < sc.n_lights && st.length >= P.min_len && bdpt_gate(P, st.length)) {
            const V3 c = mul(st.throughput, bdpt_light_radiance(sc, sc.lights[li], st, rdir));
            if (!(c.x == 0.f && c.y == 0.f && c.z == 0.f)) { emit = true; emit_c = c; }
        }
        return 0;
    }
    if (st.length >= P.max_len) return 0;
    if (bsdf.is_delta) return 2;
    if (st.length + 1 >= P.min_len && bdpt_gate(P, st.length + 1)) {     // direct illumination :205-217
        const float w = 1.f / (st.length + 1.f - st.spec);
        has_di = bdpt_direct_illumination(sc, bsdf, hit, st, st.throughput * w, di);
    }
    return 1;
}

// connectVertices (:611-665) + the gating/weighting of its caller (:231-256).  Returns true when a
// connection ray has to be traced; c.c is zero when the pair is gated out.
WRT_HD bool bdpt_connect_vertices(const DevSceneView& sc, const BdptParams& P, const LightVertex& lv, const Bsdf& cam_bsdf,
                                  V3 hit, const BdptPath& st, Connection& c)
{
    V3 dir = lv.pos - hit;
    const float dist2 = sqr_length(dir);
    const float dist = sqrtf(dist2);
    dir = div_eps(dir, dist);
    float cos_cam = 0.f, cd, cr;
    const V3 fc = bsdf_f(cam_bsdf, sc, dir, cos_cam, &cd, &cr);
    if (is_black(fc)) return false;
    cd *= cam_bsdf.continue_prob; cr *= cam_bsdf.continue_prob;
    Bsdf lb;
    bsdf_init(lb, lv.wi, lv.n, lv.matid, sc);
    float cos_l = 0.f, ld, lr;
    const V3 fl = bsdf_f(lb, sc, -dir, cos_l, &ld, &lr);
    if (is_black(fl)) return false;
    ld *= lb.continue_prob; lr *= lb.continue_prob;
    const float g = cos_l * cos_cam / dist2;
    if (cmpf(g) < 0) return false;
    const float cd_area = pdf_w_to_a(cd, dist, cos_l);
    const float ld_area = pdf_w_to_a(ld, dist, cos_cam);
    const V3 res = mul(fc, fl) * g;
    if (is_black(res)) return false;
    const bool gated_in = bdpt_gate(P, lv.length + 1 + st.length);
    if (!gated_in && !P.trace_gated) return false;
    const float w_light = cd_area * (lv.dVCM + lr * lv.dVC);
    const float w_cam = ld_area * (st.dVCM + cr * st.dVC);
    const float w = 1.f / (w_light + 1.f + w_cam);
    const float wl = 1.f / (lv.length + 1.f + st.length - lv.spec - st.spec);
    c.c = gated_in ? mul(mul(st.throughput, lv.throughput), res * w) * wl : v3(0.f, 0.f, 0.f);
    const V3 p2 = hit + dir * dist;
    c.q[0] = hit.x; c.q[1] = hit.y; c.q[2] = hit.z; c.q[3] = dir.x; c.q[4] = dir.y; c.q[5] = dir.z;
    c.q[6] = p2.x; c.q[7] = p2.y; c.q[8] = p2.z;
    c.pixel = st.index;
    return true;
}

}  // namespace wrt
