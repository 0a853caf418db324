// Whitted-style ray tracing, one call per ray-tree node: the restatement of
//   WhittedIntegrator::raytracing + directIllumination   R/src/surfaceIntegrator/whitted.cpp:17-113
//   getReflectDir / getTransDir                           R/src/math/vector.cpp:52-84
// (the next integrator on the same Scene::intersect / Scene::occluded seam.)
//
// The reference recurses: L(node) = direct*(pDiffuse+pGlossy) + pReflect*L(reflect child) + pTrans*L(trans child).
// That is linear in the children, so the tree is flattened: every node carries the product `w` of the
// probabilities on its way from the camera and adds  w * (emission | direct term)  to its pixel.  A node has
// up to two children; the wavefront keeps ONE current ray per path slot and parks the other child on a small
// per-slot pending list (at most maxTracingDepth entries), so the slot count never grows.
//
// Quirks reproduced on purpose (they shape the image): the direct term uses illu * cosAtLight / emissionPdf
// (= intensity * pi * area, no distance fall-off, :38-39) with the BSDF built around the LIGHT direction and
// evaluated towards the viewer (:35-40); a light seen from its back side gives 0 * 0 / 0 = NaN, and the reference
// adds that NaN to the pixel whenever the connection is unoccluded (:31-32 tests occlusion first); the occlusion
// query starts AT the hit point, without the EPS offset path tracing uses (:31); children are traced only while
// dep + 1 <= maxTracingDepth (:53-54).  BSDF::componentProb is read uninitialised by the reference when the hit has
// no material or the view direction grazes the surface; here those probabilities are 0.
#pragma once
#include "pt_logic.cuh"

namespace wrt {

struct WhittedChild { bool valid; RayIn ray; float w; };

struct WhittedOut {
    bool emit;            // add emit_c to the film
    bool shadow;          // occlusion query q[9] pending; add shadow_c if not occluded
    V3 emit_c, shadow_c;
    float q[9];
    WhittedChild reflect, trans;
};

WRT_HD V3 reflect_dir(V3 wi, V3 n)                       // getReflectDir
{
    const V3 res = n * dot(n, wi) * 2.0f - wi;
    return normalize(res);
}

// getTransDir; ok = false stands for the reference's Vector3(INF) result (total internal reflection)
WRT_HD V3 trans_dir(V3 wi, V3 n, float index, int inside, bool& ok)
{
    const float refraction = !inside ? index : 1.0f / index;
    const V3 N = n * (inside == 0 ? 1.0f : -1.0f);
    const float cosI = dot(N, wi);
    const float cosT = 1.0f - (refraction * refraction) * (1.0f - cosI * cosI);
    if (cmpf(cosT) > 0) {
        ok = true;
        const V3 res = (-wi) * refraction + N * (refraction * cosI - sqrtf(cosT));
        return normalize(res);
    }
    ok = false;
    return v3(WRT_INF, WRT_INF, WRT_INF);
}

WRT_HD bool v3_has_nan(V3 c) { return c.x != c.x || c.y != c.y || c.z != c.z; }

// One node of the ray tree, given the closest hit of `ray`.  `dep` is the reference's recursion depth (0 = camera
// ray), `w` the product of component probabilities so far.
WRT_HD void whitted_shade(const DevSceneView& sc, int max_depth, const RayIn& ray, int dep, float w, Rng& rng,
                          int prim, float t, WhittedOut& out)
{
    out.emit = false; out.shadow = false; out.reflect.valid = false; out.trans.valid = false;
    if (prim < 0) return;                                             // :60-61
    HitInfo h;
    fill_hit(sc, prim, ray, t, h);
    const V3 hit = v3(h.px, h.py, h.pz);
    const V3 n = v3(h.nx, h.ny, h.nz);
    const V3 rdir = v3(ray.dx, ray.dy, ray.dz);
    if (h.matid < 0) {                                                // :63-67
        const int li = -h.matid - 1;
        if (li >= sc.n_lights) return;
        out.emit = true;
        out.emit_c = light_radiance(sc.lights[li], rdir, nullptr, nullptr) * w;
        return;
    }
    Bsdf bsdf;
    bsdf_init(bsdf, -rdir, n, h.matid, sc);                           // :73

    {   // directIllumination, :17-46 (LIGHT_SAMPLE_NUM = 1; rand() % lights.size() -> our own stream)
        int k = (int)(rng_float(rng) * sc.n_lights);
        if (k >= sc.n_lights) k = sc.n_lights - 1;
        const V3 lr = rng_vec3(rng);
        V3 wi; float dist, dpdf, epdf, cosl;
        const V3 illu = light_illuminance(sc.lights[k], hit, lr, wi, dist, dpdf, &epdf, &cosl);
        Bsdf lb;
        bsdf_init(lb, wi, n, h.matid, sc);                            // :34 — built around the light direction
        float cos_wo = 0.f;
        const V3 f = bsdf_f(lb, sc, -rdir, cos_wo, nullptr, nullptr);
        const V3 e = v3(illu.x * cosl / epdf, illu.y * cosl / epdf, illu.z * cosl / epdf);   // :37-38, may be NaN
        const V3 direct = mul(e, f) * cos_wo * (bsdf.p_diffuse + bsdf.p_glossy);            // :40, :75-76
        const V3 c = direct * w;
        if (!is_black(c) || v3_has_nan(c)) {
            const V3 p2 = hit + wi * dist;
            out.q[0] = hit.x; out.q[1] = hit.y; out.q[2] = hit.z;     // occluded(inter.p, wi, inter.p + wi*dist), :31
            out.q[3] = wi.x; out.q[4] = wi.y; out.q[5] = wi.z;
            out.q[6] = p2.x; out.q[7] = p2.y; out.q[8] = p2.z;
            out.shadow_c = c;
            out.shadow = true;
        }
    }
    if (dep + 1 > max_depth) return;                                  // the children would return 0 at :53-54
    if (cmpf(bsdf.p_reflect) > 0) {                                   // :78-83
        const V3 d = reflect_dir(-rdir, n);
        make_ray(hit.x + d.x * WRT_EPS, hit.y + d.y * WRT_EPS, hit.z + d.z * WRT_EPS, d.x, d.y, d.z, out.reflect.ray);
        out.reflect.w = w * bsdf.p_reflect;
        out.reflect.valid = true;
    }
    if (cmpf(bsdf.p_trans) > 0 && bsdf.mat_id > 0 && cmpf(sc.materials[bsdf.mat_id].index) > 0) {   // :85-99
        bool ok;
        const V3 d = trans_dir(-rdir, n, sc.materials[bsdf.mat_id].index, h.inside, ok);
        if (ok && cmpf(sqr_length(d) - 1.0f) == 0) {                  // isNormal()
            make_ray(hit.x + d.x * WRT_EPS, hit.y + d.y * WRT_EPS, hit.z + d.z * WRT_EPS, d.x, d.y, d.z, out.trans.ray);
            out.trans.w = w * bsdf.p_trans;
            out.trans.valid = true;
        }
    }
}

}  // namespace wrt
