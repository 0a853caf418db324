// Path-tracing bounce logic, one call per path vertex: the restatement of
//   SurfaceIntegrator::render (per-sample body)   R/src/surfaceIntegrator/surfaceIntegrator.cpp:24-41
//   PathIntegrator::raytracing                    R/src/surfaceIntegrator/pathIntegrator.cpp:29-148
// split at the two places the reference calls into the scene (scene.intersect :45, scene.occluded
// :95) so that the traversal runs as its own kernels (wavefront).  RNG draw order per vertex is
// the reference's: [1 + 3] NEE if non-delta, 3 BSDF, [1] Russian roulette.
#pragma once
#include "shading.cuh"

namespace wrt {

struct PtParams {
    int width, height, spp, strata, max_depth;
    uint32_t seed;
    int sample_first, sample_stride, local_spp;   // this call renders local_spp samples per pixel
    float film_scale;
    unsigned long long total_samples;             // width*height*local_spp
    uint32_t tape_stride;                         // > 0: RNG replay, sample (pixel, k) reads tape[(pixel*spp + k)*stride ...]
};

struct PathData {      // per-path state besides the ray
    V3 weight;         // pathWeight
    float last_pdf;
    uint32_t pixel;    // i*width + j
    Rng rng;
    int length;        // pathLength
    int last_specular;
};

struct ShadeOut {
    bool alive;        // next ray is in `ray`
    bool emit;         // add `emit_c` to the film (emitter hit)
    bool shadow;       // shadow query q[9] pending; add shadow_c if not occluded
    V3 emit_c, shadow_c;
    float q[9];
};

// Sample k of pixel (row i, column j): sampleRectangleStratified over [j-1/2, j+1/2] x [i-1/2, i+1/2] on a
// strata x strata grid (sampler.cpp:28-42; v0 + p1*a + p2*b with p1 = (1,0,0), p2 = (0,1,0)) + Camera::generateRay.
WRT_HD void pt_sample_ray(int strata, const DevCamera& cam, int i, int j, int k, float ux, float uy, RayIn& ray)
{
    const int len = strata;
    const int row = k / len, col = k % len;
    const float a = (ux + row) / (float)len;
    const float b = (uy + col) / (float)len;
    const float x = (j - 0.5f) + 1.f * a + 0.f * b;
    const float y = (i - 0.5f) + 0.f * a + 1.f * b;
    camera_ray(cam, x, y, ray);
}

// One camera sample: stratified jitter (sampler.cpp:28-42) + Camera::generateRay.
WRT_HD void pt_generate(const PtParams& P, const DevCamera& cam, unsigned long long s, RayIn& ray, PathData& pd)
{
    const unsigned long long npix = (unsigned long long)P.width * P.height;
    const uint32_t lin = (uint32_t)(s % npix);
    const int k = P.sample_first + (int)(s / npix) * P.sample_stride;
    // Consecutive samples walk the image in 8x4 pixel tiles (one warp = one tile) when the size allows:
    // neighbouring rays share nodes and leaves.  The RNG is keyed on the pixel, so the image is unchanged.
    int i, j;
    if ((P.width & 7) == 0 && (P.height & 3) == 0) {
        const uint32_t tile = lin >> 5, within = lin & 31u, tiles_x = (uint32_t)P.width >> 3;
        j = (int)((tile % tiles_x) * 8u + (within & 7u));
        i = (int)((tile / tiles_x) * 4u + (within >> 3));
    } else {
        i = (int)(lin / (uint32_t)P.width); j = (int)(lin % (uint32_t)P.width);
    }
    const uint32_t pixel = (uint32_t)i * (uint32_t)P.width + (uint32_t)j;
    pd.rng = P.tape_stride ? rng_make_tape((unsigned long long)pixel * (unsigned)P.spp + (unsigned)k, P.tape_stride)
                           : rng_make(P.seed, 0u, (unsigned long long)k * npix + pixel);
    const V3 u = rng_vec3(pd.rng);
    pt_sample_ray(P.strata, cam, i, j, k, u.x, u.y, ray);
    pd.weight = v3(1.f, 1.f, 1.f);
    pd.last_pdf = 1.f;
    pd.pixel = pixel;
    pd.length = 1;
    pd.last_specular = 1;
}

// One vertex of PathIntegrator::raytracing's loop body, given the closest hit of `ray`.
WRT_HD void pt_shade(const DevSceneView& sc, const PtParams& P, RayIn& ray, PathData& pd, int prim, float t,
                     ShadeOut& out)
{
    out.alive = false; out.emit = false; out.shadow = false;
    if (prim < 0) return;                                             // :45-46
    HitInfo h;
    fill_hit(sc, prim, ray, t, h);
    const V3 hit = v3(h.px, h.py, h.pz);
    const V3 rdir = v3(ray.dx, ray.dy, ray.dz);
    Bsdf bsdf;
    bsdf_init(bsdf, -rdir, v3(h.nx, h.ny, h.nz), h.matid, sc);
    if (bsdf.mat_id == 0) return;                                     // :50-51
    const float pick = 1.f / sc.n_lights;                             // lightPickProb :39

    if (bsdf.mat_id < 0) {                                            // hit light directly :54-73
        const int li = -bsdf.mat_id - 1;
        if (li >= sc.n_lights) return;
        float pdf_area;
        const V3 c = light_radiance(sc.lights[li], rdir, &pdf_area, nullptr);
        if (is_black(c)) return;
        float w = 1.f;
        if (pd.length > 1 && !pd.last_specular) {
            const float direct_pdf = pdf_a_to_w(pdf_area, t, bsdf.wi_local.z);
            w = pd.last_pdf / (pd.last_pdf + direct_pdf * pick);
        }
        out.emit = true;
        out.emit_c = mul(pd.weight, c) * w;
        return;
    }
    if (pd.length > P.max_depth) return;                              // :75-76
    if (cmpf(bsdf.continue_prob) == 0) return;                        // :78-79

    if (!bsdf.is_delta) {                                             // direct illumination :82-118
        int light_id = (int)(rng_float(pd.rng) * sc.n_lights);
        if (light_id >= sc.n_lights) light_id = sc.n_lights - 1;
        const V3 lr = rng_vec3(pd.rng);
        V3 dl; float dist, dpdf;
        const V3 illu = light_illuminance(sc.lights[light_id], hit, lr, dl, dist, dpdf, nullptr, nullptr);
        if (!is_black(illu)) {
            float bpdf, cos_wo = 0.f;
            const V3 f = bsdf_f(bsdf, sc, dl, cos_wo, &bpdf, nullptr);
            // The reference traces the shadow ray first and evaluates f afterwards (:94-99); a black
            // f makes the ray's outcome irrelevant, so the ray is only queued when it can contribute.
            if (!is_black(f)) {
                const V3 p1 = hit + dl * WRT_EPS;
                const V3 p2 = hit + dl * (dist - WRT_EPS);
                out.q[0] = p1.x; out.q[1] = p1.y; out.q[2] = p1.z;
                out.q[3] = dl.x; out.q[4] = dl.y; out.q[5] = dl.z;
                out.q[6] = p2.x; out.q[7] = p2.y; out.q[8] = p2.z;
                bpdf *= bsdf.continue_prob;
                const float w = (dpdf * pick) / (dpdf * pick + bpdf);
                const V3 contrib = mul(illu, f) * (w * cos_wo / (pick * dpdf));
                out.shadow_c = mul(contrib, pd.weight);
                out.shadow = true;
            }
        }
    }

    float pdf = 0.f, cos_wo = 0.f; int type = 0;                      // continue random walk :120-146
    V3 wo = rdir;
    const V3 f = bsdf_sample(bsdf, sc, rng_vec3(pd.rng), wo, pdf, cos_wo, type);
    if (is_black(f)) return;
    const float cont = bsdf.continue_prob;
    pd.last_specular = (type & BSDF_SPECULAR) != 0;
    pd.last_pdf = pdf * cont;
    if (cmpf(cont - 1.f) < 0) {
        if (cmpf(rng_float(pd.rng) - cont) > 0) return;
        pdf *= cont;
    }
    pd.weight = mul(pd.weight, f) * (cos_wo / pdf);
    ray.ox = hit.x + wo.x * WRT_EPS; ray.oy = hit.y + wo.y * WRT_EPS; ray.oz = hit.z + wo.z * WRT_EPS;
    ray.dx = wo.x; ray.dy = wo.y; ray.dz = wo.z;                      // r.dir is NOT re-normalised (:128,144)
    ray.tmin = 0.f; ray.tmax = WRT_INF;
    pd.length += 1;
    out.alive = true;
}

}  // namespace wrt
