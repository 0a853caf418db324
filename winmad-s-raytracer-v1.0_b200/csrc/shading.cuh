// Shading math the bounce loops call, restated as __host__ __device__ functions:
//   Vector3 / Color3 / cmp / clampVal      R/src/math/{vector,color,math}.{h,cpp}
//   Frame                                   R/src/math/frame.cpp:3-21
//   samplers                                R/src/sampler/sampler.cpp:3-16,28-42,95-135
//   fresnelDielectric                       R/src/material/fresnel.cpp:3-30
//   BSDF (init/f/pdf/sample)                R/src/material/bsdf.h:66-94, bsdf.cpp:24-335
//   AreaLight (illuminance/emit/getRadiance) R/src/scene/light.cpp:4-100
//   Camera::generateRay / Transform::tPoint R/src/scene/camera.cpp:37-42, math/transform.h:126-140
// plus the counter-based RNG that replaces the reference's single MT19937 stream (rng.cpp); it
// keeps the reference's float convention (u32 & 0xffffff) / 2^24 (rng.cpp:18-22).
// Image parity with the reference is statistical (cosf/sinf/powf differ by ulps, RNG differs by
// construction); the reference's quirks that bias the image are kept: isBlack() treats channels
// within 1e-3 of zero as zero (color.h:31-35), material 0 kills paths (bsdf.h:91-94), emission is
// one sided (light.cpp:78-88).
#pragma once
#include "traverse.cuh"

namespace wrt {

struct V3 { float x, y, z; };

WRT_HD V3 v3(float x, float y, float z) { V3 r = { x, y, z }; return r; }
WRT_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
WRT_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
WRT_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
WRT_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
WRT_HD V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }           // operator|
WRT_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            // operator^
WRT_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
WRT_HD float sqr_length(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
WRT_HD float length(V3 a) { return sqrtf(sqr_length(a)); }
WRT_HD V3 normalize(V3 a) { float len = sqrtf(sqr_length(a)); return v3(a.x / len, a.y / len, a.z / len); }
WRT_HD int cmpf(float x) { return (x < -WRT_EPS) ? -1 : (x > WRT_EPS ? 1 : 0); }
// Vector3 operator/ (vector.cpp:35-39): INF vector when |right| <= EPS
WRT_HD V3 div_eps(V3 a, float s) { if (cmpf(s) == 0) return v3(WRT_INF, WRT_INF, WRT_INF); return v3(a.x / s, a.y / s, a.z / s); }
WRT_HD float clamp_val(float v, float lo, float hi)   // std::min(hi, std::max(v, lo))
{
    float m = (v < lo) ? lo : v;
    return (m < hi) ? m : hi;
}
WRT_HD bool is_black(V3 c) { return cmpf(c.x) == 0 && cmpf(c.y) == 0 && cmpf(c.z) == 0; }
WRT_HD float luminance(const float c[3]) { return 0.2126f * c[0] + 0.7152f * c[1] + 0.0722f * c[2]; }
WRT_HD V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }

#define WRT_PI 3.14159274101257324f       /* acos(-1.0f) */
#define WRT_INV_PI (1.0f / WRT_PI)

WRT_HD float pdf_w_to_a(float pdfW, float dist, float c) { return pdfW * fabsf(c) / (dist * dist); }   // math.cpp:13-16
WRT_HD float pdf_a_to_w(float pdfA, float dist, float c) { return pdfA * (dist * dist) / fabsf(c); }   // math.cpp:18-21

// ---- RNG ------------------------------------------------------------------------------------------
WRT_HD uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
struct Rng { uint32_t key, ctr; };
WRT_HD Rng rng_make(uint32_t seed, uint32_t stream, uint64_t sample_id)
{
    Rng r;
    uint32_t h = hash32(seed ^ 0x9E3779B9u);
    h = hash32(h ^ (uint32_t)sample_id);
    h = hash32(h + (uint32_t)(sample_id >> 32) * 0x85EBCA6Bu + stream * 0xC2B2AE35u);
    r.key = h; r.ctr = 0;
    return r;
}
WRT_HD uint32_t rng_u32(Rng& r) { uint32_t c = r.ctr++; return hash32(r.key ^ hash32(c * 0x9E3779B9u + 0x6A09E667u)); }

// Replay mode (diagnostics, wrt_debug_set_rng_tape): when a tape is installed, stream `key` reads its numbers from
// tape[key + ctr] instead of hashing — the tape holds, per sample, the floats the reference's MT19937 stream delivered
// to that sample, so the device follows the reference's paths sample for sample (tests/test_*tape*).
#if defined(__CUDACC__)
static __constant__ const float* c_rng_tape = nullptr;    // one copy per translation unit, set by its own entry points
#endif
static const float* h_rng_tape = nullptr;                 // host builds of this code (tests/hostsim)
WRT_HD const float* rng_tape()
{
#ifdef __CUDA_ARCH__
    return c_rng_tape;
#else
    return h_rng_tape;
#endif
}
WRT_HD Rng rng_make_tape(uint64_t sample_index, uint32_t stride) { Rng r; r.key = (uint32_t)(sample_index * stride); r.ctr = 0; return r; }
WRT_HD float rng_float(Rng& r)
{
    const float* tape = rng_tape();
    if (tape) return tape[(size_t)r.key + r.ctr++];
    return (float)(rng_u32(r) & 0xffffffu) / 16777216.f;
}
WRT_HD V3 rng_vec3(Rng& r) { float a = rng_float(r), b = rng_float(r), c = rng_float(r); return v3(a, b, c); }

// ---- Frame ------------------------------------------------------------------------------------------
struct Frame { V3 x, y, z; };
WRT_HD Frame frame_from_z(V3 zin)
{
    Frame f;
    f.z = normalize(zin);
    V3 tmpx = (fabsf(f.z.x) > 0.99f) ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
    f.y = normalize(cross(f.z, tmpx));
    f.x = cross(f.y, f.z);
    return f;
}
WRT_HD V3 to_world(const Frame& f, V3 l) { return f.x * l.x + f.y * l.y + f.z * l.z; }
WRT_HD V3 to_local(const Frame& f, V3 w) { return v3(dot(w, f.x), dot(w, f.y), dot(w, f.z)); }

// ---- samplers ---------------------------------------------------------------------------------------
WRT_HD V3 sample_triangle(V3 s, V3 v1, V3 v2, V3 v3_)
{
    V3 p1 = v2 - v1, p2 = v3_ - v1;
    float u1 = sqrtf(s.x);
    float beta = 1.f - u1, gamma = s.y * u1;
    return v1 + p1 * beta + p2 * gamma;
}
WRT_HD V3 sample_cos_hemisphere(V3 s, float* pdf)
{
    float u1 = 2.f * WRT_PI * s.x;
    float u2 = sqrtf(1.f - s.y);
    V3 res = v3(cosf(u1) * u2, sinf(u1) * u2, sqrtf(s.y));
    if (pdf) *pdf = res.z * WRT_INV_PI;
    return normalize(res);
}
WRT_HD float cos_hemisphere_pdf(V3 n, V3 dir) { return clamp_val(dot(n, dir), 0.f, 1.f) * WRT_INV_PI; }
WRT_HD V3 sample_power_cos_hemisphere(V3 s, float power)
{
    float u1 = 2.f * WRT_PI * s.x;
    float u2 = powf(s.y, 1.f / (power + 1.f));
    float u3 = sqrtf(1.f - u2 * u2);
    return normalize(v3(cosf(u1) * u3, sinf(u1) * u3, u2));
}
WRT_HD float power_cos_hemisphere_pdf(V3 n, V3 dir, float power)
{
    float c = clamp_val(dot(n, dir), 0.f, 1.f);
    return (power + 1.f) * powf(c, power) * (0.5f * WRT_INV_PI);
}

// ---- Fresnel ----------------------------------------------------------------------------------------
WRT_HD float fresnel_dielectric(float cosI, float index)
{
    if (cmpf(index) < 0) return 1.0f;
    float eta;
    if (cmpf(cosI) < 0) { cosI = -cosI; eta = index; } else eta = 1.0f / index;
    float sinT2 = (eta * eta) * (1.0f - cosI * cosI);
    float a = 1.0f - sinT2;
    float cosT = sqrtf((0.0f < a) ? a : 0.0f);    // std::max(0.0f, a)
    float term1 = eta * cosT;
    float parallel = (cosI - term1) / (cosI + term1);
    float term2 = eta * cosI;
    float perpendicular = (term2 - cosT) / (term2 + cosT);
    return 0.5f * (parallel * parallel + perpendicular * perpendicular);
}

// ---- BSDF -------------------------------------------------------------------------------------------
enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 3 };

struct Bsdf {
    int mat_id;            // 0 = invalid
    Frame frame;
    V3 wi_local;
    bool is_delta;
    float continue_prob, fresnel_reflect;
    float p_diffuse, p_glossy, p_reflect, p_trans;
};

// BSDF::init (bsdf.h:72-89) + calcComponentProb (bsdf.cpp:24-55)
WRT_HD void bsdf_init(Bsdf& b, V3 wi, V3 n, int inter_mat, const DevSceneView& sc)
{
    b.mat_id = 0;
    b.is_delta = false; b.continue_prob = 0.f; b.fresnel_reflect = 0.f;
    b.p_diffuse = b.p_glossy = b.p_reflect = b.p_trans = 0.f;
    b.frame = frame_from_z(n);
    b.wi_local = normalize(to_local(b.frame, wi));
    if (cmpf(b.wi_local.z) == 0) return;
    if (inter_mat > 0 && inter_mat < sc.n_materials) {
        const DevMaterial& m = sc.materials[inter_mat];
        b.fresnel_reflect = fresnel_dielectric(b.wi_local.z, m.index);
        float pd = luminance(m.diffuse), pg = luminance(m.phong);
        float pr = b.fresnel_reflect * luminance(m.specular);
        float pt = (1.f - b.fresnel_reflect) * 1.0f;
        float total = pd + pg + pr + pt;
        if (cmpf(total) <= 0) {
            b.p_diffuse = b.p_glossy = b.p_reflect = b.p_trans = 0.f; b.continue_prob = 0.f;
        } else {
            b.p_diffuse = pd / total; b.p_glossy = pg / total; b.p_reflect = pr / total; b.p_trans = pt / total;
            float r0 = m.diffuse[0] + m.phong[0] + m.specular[0] * b.fresnel_reflect;
            float r1 = m.diffuse[1] + m.phong[1] + m.specular[1] * b.fresnel_reflect;
            float r2 = m.diffuse[2] + m.phong[2] + m.specular[2] * b.fresnel_reflect;
            float mx12 = (r1 < r2) ? r2 : r1;                 // std::max(g, b)
            float mx = (r0 < mx12) ? mx12 : r0;               // std::max(r, ...)
            b.continue_prob = clamp_val(mx + (1.f - b.fresnel_reflect), 0.f, 1.f);
        }
    } else if (inter_mat > 0) {
        return;  // material index past the table: the reference reads out of bounds; we treat it as material 0
    }
    b.is_delta = (cmpf(b.p_diffuse) == 0 && cmpf(b.p_glossy) == 0);
    b.mat_id = inter_mat;
}

WRT_HD V3 bsdf_calc_diffuse(const Bsdf& b, const DevMaterial& m, V3 wo, float* dpdf, float* rpdf)
{
    if (cmpf(b.p_diffuse) == 0) return v3(0, 0, 0);
    if (cmpf(b.wi_local.z) <= 0 || cmpf(wo.z) <= 0) return v3(0, 0, 0);
    if (dpdf) *dpdf += b.p_diffuse * clamp_val(wo.z * WRT_INV_PI, 0.0f, 1.0f);
    if (rpdf) *rpdf += b.p_diffuse * clamp_val(b.wi_local.z * WRT_INV_PI, 0.0f, 1.0f);
    return ld3(m.diffuse) * WRT_INV_PI;
}

WRT_HD V3 bsdf_calc_glossy(const Bsdf& b, const DevMaterial& m, V3 wo, float* dpdf, float* rpdf)
{
    if (cmpf(b.p_glossy) == 0) return v3(0, 0, 0);
    if (cmpf(b.wi_local.z) <= 0 || cmpf(wo.z) <= 0) return v3(0, 0, 0);
    V3 refl = v3(-b.wi_local.x, -b.wi_local.y, b.wi_local.z);
    float c = dot(refl, wo);
    if (cmpf(c) == 0) return v3(0, 0, 0);
    float pdfW = b.p_glossy * power_cos_hemisphere_pdf(refl, wo, m.phong_exp);
    if (dpdf) *dpdf += pdfW;
    if (rpdf) *rpdf += pdfW;
    V3 rho = ld3(m.phong) * (m.phong_exp + 2.f) * 0.5f * WRT_INV_PI;
    return rho * powf(c, m.phong_exp);
}

// BSDF::f (bsdf.cpp:102-126)
WRT_HD V3 bsdf_f(const Bsdf& b, const DevSceneView& sc, V3 wo_world, float& cos_wo, float* dpdf, float* rpdf)
{
    if (dpdf) *dpdf = 0.f;
    if (rpdf) *rpdf = 0.f;
    V3 wo = to_local(b.frame, wo_world);
    if (cmpf(wo.z * b.wi_local.z) < 0) return v3(0, 0, 0);
    cos_wo = fabsf(wo.z);
    if (b.mat_id < 0) return v3(0, 0, 0);
    const DevMaterial& m = sc.materials[b.mat_id];
    V3 res = v3(0, 0, 0);
    res = res + bsdf_calc_diffuse(b, m, wo, dpdf, rpdf);
    res = res + bsdf_calc_glossy(b, m, wo, dpdf, rpdf);
    return res;
}

WRT_HD void bsdf_pdf_glossy(const Bsdf& b, const DevMaterial& m, V3 wo, float* dpdf, float* rpdf)
{
    if (cmpf(b.p_glossy) == 0) return;
    V3 refl = v3(-b.wi_local.x, -b.wi_local.y, b.wi_local.z);
    float c = dot(refl, wo);
    if (cmpf(c) == 0) return;
    float pdfW = b.p_glossy * power_cos_hemisphere_pdf(refl, wo, m.phong_exp);
    if (dpdf) *dpdf += pdfW;
    if (rpdf) *rpdf += pdfW;
}

// BSDF::pdf (bsdf.cpp:166-183)
WRT_HD float bsdf_pdf(const Bsdf& b, const DevSceneView& sc, V3 wo_world, bool rev)
{
    V3 wo = to_local(b.frame, wo_world);
    if (cmpf(wo.z * b.wi_local.z) < 0) return 0.f;
    const DevMaterial& m = sc.materials[b.mat_id];
    float d = 0.f, r = 0.f;
    if (cmpf(b.p_diffuse) != 0) {                                   // pdfDiffuse, bsdf.cpp:128-142
        d += b.p_diffuse * clamp_val(wo.z, 0.f, 1.f) * WRT_INV_PI;
        r += b.p_diffuse * clamp_val(b.wi_local.z, 0.f, 1.f) * WRT_INV_PI;
    }
    bsdf_pdf_glossy(b, m, wo, &d, &r);
    return rev ? r : d;
}

// BSDF::sample (bsdf.cpp:268-335).  Returns the BSDF factor; black means "terminate".
WRT_HD V3 bsdf_sample(const Bsdf& b, const DevSceneView& sc, V3 rnd, V3& wo_world, float& pdf, float& cos_wo, int& type)
{
    if (rnd.z < b.p_diffuse) type = BSDF_DIFFUSE;
    else if (rnd.z < b.p_diffuse + b.p_glossy) type = BSDF_GLOSSY;
    else if (rnd.z < b.p_diffuse + b.p_glossy + b.p_reflect) type = BSDF_REFLECTION;
    else type = BSDF_TRANSMISSION;
    if (b.mat_id < 0) return v3(0, 0, 0);
    const DevMaterial& m = sc.materials[b.mat_id];
    pdf = 0.f;
    V3 res = v3(0, 0, 0);
    V3 wo = v3(0, 0, 0);
    if (type == BSDF_DIFFUSE) {
        if (cmpf(b.wi_local.z) <= 0) return v3(0, 0, 0);           // sampleDiffuse :185-197
        float pw;
        wo = sample_cos_hemisphere(rnd, &pw);
        pdf += pw * b.p_diffuse;
        res = res + ld3(m.diffuse) * WRT_INV_PI;
        if (is_black(res)) return v3(0, 0, 0);
        res = res + bsdf_calc_glossy(b, m, wo, &pdf, nullptr);
    } else if (type == BSDF_GLOSSY) {                               // sampleGlossy :199-218
        wo = sample_power_cos_hemisphere(rnd, m.phong_exp);
        V3 refl = v3(-b.wi_local.x, -b.wi_local.y, b.wi_local.z);
        Frame fr = frame_from_z(refl);
        wo = to_world(fr, wo);
        float c = dot(refl, wo);
        if (cmpf(c) <= 0) return v3(0, 0, 0);
        bsdf_pdf_glossy(b, m, wo, &pdf, nullptr);
        V3 rho = ld3(m.phong) * (m.phong_exp + 2.f) * 0.5f * WRT_INV_PI;
        res = res + rho * powf(c, m.phong_exp);
        if (is_black(res)) return v3(0, 0, 0);
        res = res + bsdf_calc_diffuse(b, m, wo, &pdf, nullptr);
    } else if (type == BSDF_REFLECTION) {                           // sampleReflect :220-226
        wo = v3(-b.wi_local.x, -b.wi_local.y, b.wi_local.z);
        pdf += b.p_reflect;
        // mat.specular * fresnelReflect / abs(z)  (Color3 * Real, then Color3 / Real per channel)
        res = v3(m.specular[0] * b.fresnel_reflect / fabsf(wo.z), m.specular[1] * b.fresnel_reflect / fabsf(wo.z),
                 m.specular[2] * b.fresnel_reflect / fabsf(wo.z));
        if (is_black(res)) return v3(0, 0, 0);
    } else {                                                        // sampleTrans :228-266
        if (cmpf(m.index) < 0) return v3(0, 0, 0);
        float cosI = b.wi_local.z, cosT, eta;
        if (cmpf(cosI) < 0) { eta = m.index; cosI = -cosI; cosT = 1.f; }
        else { eta = 1.f / m.index; cosT = -1.f; }
        float sinI2 = 1.f - cosI * cosI;
        float sinT2 = (eta * eta) * sinI2;
        if (sinT2 < 1.f) {
            cosT *= sqrtf(clamp_val(1.f - sinT2, 0.f, 1.f));
            wo = normalize(v3(-eta * b.wi_local.x, -eta * b.wi_local.y, cosT));
            pdf += b.p_trans;
            float tc = 1.f - b.fresnel_reflect;
            float v = tc / fabsf(cosT);
            res = v3(v, v, v);
        } else return v3(0, 0, 0);
        if (is_black(res)) return v3(0, 0, 0);
    }
    cos_wo = fabsf(wo.z);
    if (cmpf(cos_wo) == 0) return v3(0, 0, 0);
    wo_world = to_world(b.frame, wo);
    return res;
}

// ---- AreaLight --------------------------------------------------------------------------------------
WRT_HD V3 light_normal(const DevLight& L) { return ld3(L.fz); }

// AreaLight::illuminance (light.cpp:4-38)
WRT_HD V3 light_illuminance(const DevLight& L, V3 pos, V3 rnd, V3& dir_to_light, float& dist, float& direct_pdf,
                            float* emission_pdf, float* cos_at_light)
{
    if (emission_pdf) *emission_pdf = 0.f;
    if (cos_at_light) *cos_at_light = 0.f;
    V3 p0 = ld3(L.p0);
    V3 lp = sample_triangle(rnd, p0, p0 + ld3(L.d1), p0 + ld3(L.d2));
    dir_to_light = lp - pos;
    dist = length(dir_to_light);
    dir_to_light = div_eps(dir_to_light, dist);
    float cos_n = dot(light_normal(L), -dir_to_light);
    if (cmpf(cos_n) <= 0) { direct_pdf = 0.f; return v3(0, 0, 0); }
    direct_pdf = L.inv_area * (dist * dist) / cos_n;
    if (cos_at_light) *cos_at_light = cos_n;
    if (emission_pdf) *emission_pdf = L.inv_area * cos_n * WRT_INV_PI;
    return ld3(L.intensity);
}

// AreaLight::emit (light.cpp:40-67)
WRT_HD V3 light_emit(const DevLight& L, V3 dir_rnd, V3 pos_rnd, V3& pos, V3& dir, float& emission_pdf,
                     float& direct_pdf_area, float& cos_at_light)
{
    V3 p0 = ld3(L.p0);
    pos = sample_triangle(pos_rnd, p0, p0 + ld3(L.d1), p0 + ld3(L.d2));
    V3 local = sample_cos_hemisphere(dir_rnd, &emission_pdf);
    emission_pdf *= L.inv_area;
    local.z = (local.z < WRT_EPS) ? WRT_EPS : local.z;            // std::max(z, EPS)
    Frame f; f.x = ld3(L.fx); f.y = ld3(L.fy); f.z = ld3(L.fz);
    dir = to_world(f, local);
    direct_pdf_area = L.inv_area;
    cos_at_light = local.z;
    return ld3(L.intensity) * local.z;
}

// AreaLight::getRadiance (light.cpp:69-100)
WRT_HD V3 light_radiance(const DevLight& L, V3 ray_dir, float* direct_pdf_area, float* emission_pdf)
{
    if (direct_pdf_area) *direct_pdf_area = 0.f;
    if (emission_pdf) *emission_pdf = 0.f;
    float cos_n = clamp_val(dot(light_normal(L), -ray_dir), 0.f, 1.f);
    if (cmpf(cos_n) == 0) return v3(0, 0, 0);
    if (direct_pdf_area) *direct_pdf_area = L.inv_area;
    if (emission_pdf) *emission_pdf = cos_hemisphere_pdf(light_normal(L), -ray_dir) * L.inv_area;
    return ld3(L.intensity);
}

// ---- camera -----------------------------------------------------------------------------------------
struct DevCamera {
    float pos[3], forward[3];
    float image_plane_dist, x_res, y_res;
    float r2w[16], w2r[16];
};

WRT_HD V3 t_point(const float* m, V3 p)   // Transform::tPoint
{
    float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (cmpf(wp - 1.0f) == 0) return v3(xp, yp, zp);
    return div_eps(v3(xp, yp, zp), wp);
}

WRT_HD void camera_ray(const DevCamera& c, float x, float y, RayIn& r)   // Camera::generateRay
{
    V3 p = t_point(c.r2w, v3(x, y, 0.f));
    make_ray(c.pos[0], c.pos[1], c.pos[2], p.x - c.pos[0], p.y - c.pos[1], p.z - c.pos[2], r);
}

WRT_HD bool camera_check_raster(const DevCamera& c, float x, float y)   // Camera::checkRaster
{
    return cmpf(x) >= 0 && cmpf(y) >= 0 && cmpf(x - c.x_res) < 0 && cmpf(y - c.y_res) < 0;
}

}  // namespace wrt
