// Known-answer entry of the shading math (diagnostics): evaluates ONE of the restated reference functions on one input
// record.  The layout of `in` / `out` per `what` is the one oracle/ref_harness.cpp::ref_shading_batch uses for the
// reference's own functions, so the device (wrt_debug_shading), the host build (tests/hostsim) and the reference can be
// compared entry by entry:
//   0 BSDF::f (bsdf.cpp:102-126)             1 BSDF::sample (:268-335)         2 BSDF::pdf (:166-183)
//   3 AreaLight::illuminance (light.cpp:4-38) 4 AreaLight::emit (:40-67)        5 AreaLight::getRadiance (:69-100)
//   6 fresnelDielectric (fresnel.cpp:3-30)    7 samplers (sampler.cpp:3-16,95-135)
//   8 camera sample: sampleRectangleStratified + Camera::generateRay (surfaceIntegrator.cpp:26-34)
#pragma once
#include "pt_logic.cuh"

namespace wrt {

WRT_HD int shading_kat_in_stride(int what)
{
    return what == 0 || what == 1 || what == 2 ? 10 : (what == 3 || what == 4 ? 7 : (what == 5 ? 4 : (what == 6 ? 2 : (what == 7 ? 13 : (what == 8 ? 5 : 0)))));
}
WRT_HD int shading_kat_out_stride(int what)
{
    return what == 0 || what == 1 ? 9 : (what == 2 ? 2 : (what == 3 ? 10 : (what == 4 ? 12 : (what == 5 ? 5 : (what == 6 ? 1 : (what == 7 ? 13 : (what == 8 ? 8 : 0)))))));
}

WRT_HD void shading_kat(const DevSceneView& sc, const DevCamera& cam, int what, int iparam, const float* a, float* o)
{
    const int os = shading_kat_out_stride(what);
    for (int k = 0; k < os; k++) o[k] = 0.f;
    if (what <= 2) {
        const int matid = (int)a[6];
        Bsdf b;
        bsdf_init(b, v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), matid, sc);
        const bool valid = b.mat_id != 0;
        if (what == 0) {
            o[8] = valid ? 1.f : 0.f;
            if (!valid || matid <= 0) return;
            float cos_wo = 0.f, dp = 0.f, rp = 0.f;
            const V3 f = bsdf_f(b, sc, v3(a[7], a[8], a[9]), cos_wo, &dp, &rp);
            o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = cos_wo; o[4] = dp; o[5] = rp; o[6] = b.continue_prob; o[7] = b.is_delta ? 1.f : 0.f;
        } else if (what == 1) {
            if (!valid || matid <= 0) return;
            V3 wo = v3(0.f, 0.f, 0.f); float pdf = 0.f, cos_wo = 0.f; int type = 0;
            const V3 f = bsdf_sample(b, sc, v3(a[7], a[8], a[9]), wo, pdf, cos_wo, type);
            o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = wo.x; o[4] = wo.y; o[5] = wo.z; o[6] = pdf; o[7] = cos_wo; o[8] = (float)type;
        } else {
            if (!valid || matid <= 0) return;
            o[0] = bsdf_pdf(b, sc, v3(a[7], a[8], a[9]), false);
            o[1] = bsdf_pdf(b, sc, v3(a[7], a[8], a[9]), true);
        }
    } else if (what == 3) {
        V3 d = v3(0.f, 0.f, 0.f); float dist = 0.f, dpdf = 0.f, epdf = 0.f, cosl = 0.f;
        const V3 c = light_illuminance(sc.lights[(int)a[0]], v3(a[1], a[2], a[3]), v3(a[4], a[5], a[6]), d, dist, dpdf, &epdf, &cosl);
        o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = d.x; o[4] = d.y; o[5] = d.z; o[6] = dist; o[7] = dpdf; o[8] = epdf; o[9] = cosl;
    } else if (what == 4) {
        V3 p = v3(0.f, 0.f, 0.f), d = v3(0.f, 0.f, 0.f); float epdf = 0.f, dpdf = 0.f, cosl = 0.f;
        const V3 c = light_emit(sc.lights[(int)a[0]], v3(a[1], a[2], a[3]), v3(a[4], a[5], a[6]), p, d, epdf, dpdf, cosl);
        o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = p.x; o[4] = p.y; o[5] = p.z; o[6] = d.x; o[7] = d.y; o[8] = d.z; o[9] = epdf; o[10] = dpdf; o[11] = cosl;
    } else if (what == 5) {
        float dpa = 0.f, epdf = 0.f;
        const V3 c = light_radiance(sc.lights[(int)a[0]], v3(a[1], a[2], a[3]), &dpa, &epdf);
        o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = dpa; o[4] = epdf;
    } else if (what == 6) {
        o[0] = fresnel_dielectric(a[0], a[1]);
    } else if (what == 7) {
        const V3 r = v3(a[0], a[1], a[2]);
        const V3 p = sample_triangle(r, v3(a[4], a[5], a[6]), v3(a[7], a[8], a[9]), v3(a[10], a[11], a[12]));
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
        float pdf = 0.f;
        const V3 c = sample_cos_hemisphere(r, &pdf);
        o[3] = c.x; o[4] = c.y; o[5] = c.z; o[6] = pdf;
        const V3 g = sample_power_cos_hemisphere(r, a[3]);
        o[7] = g.x; o[8] = g.y; o[9] = g.z;
        o[10] = cos_hemisphere_pdf(v3(0.f, 0.f, 1.f), c);
        o[11] = power_cos_hemisphere_pdf(v3(0.f, 0.f, 1.f), g, a[3]);
    } else if (what == 8) {
        int strata = (int)sqrt((double)iparam);
        if (strata < 1) strata = 1;
        RayIn r;
        pt_sample_ray(strata, cam, (int)a[2], (int)a[3], (int)a[4], a[0], a[1], r);
        o[0] = r.ox; o[1] = r.oy; o[2] = r.oz; o[3] = r.dx; o[4] = r.dy; o[5] = r.dz; o[6] = r.tmin; o[7] = r.tmax;
    }
}

}  // namespace wrt
