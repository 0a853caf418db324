// KD-tree traversal + ray/triangle + ray/sphere intersection: the device restatement of
//   KDtreeAccel::traverse   R/src/scene/KDtreeAccel.cpp:309-388
//   AABB::hit               R/src/geometry/AABB.cpp:9-32
//   Triangle::hit           R/src/geometry/triangle.cpp:22-87
//   Sphere::hit             R/src/geometry/sphere.cpp:17-78
//   cmp()                   R/src/math/math.cpp:8-11 (EPS = 1e-3f)
//
// Result contract: the winning primitive id and its t are bit-identical to the reference's.  That
// needs (a) the reference's own tree (visit order), (b) its acceptance rule `t - best < -EPS`
// applied in visit order, (c) IEEE float arithmetic in the reference's operand order with NO fused
// multiply-add: this translation unit is compiled with -fmad=false, and divisions/square roots are
// the IEEE-rounded ones (nvcc default -prec-div=true -prec-sqrt=true, no -use_fast_math).
// std::max/std::min/std::swap are spelled out as the comparisons they perform so that NaNs (0*inf
// from axis-parallel rays) take the same branches as on the CPU.
//
// PRUNED mode keeps the visit order but skips a sub-tree — or a chunk of a leaf's records (skip records) —
// when a conservative bound of everything referenced in it (built by scene_layout.cpp) is entered by the ray
// later than best*(1+1e-4), or is not entered at all: no hit there can satisfy `t < best - EPS` / exist, so
// skipping never changes the state the reference would have (box_prunable).  tests/test_gpu_traversal.py and
// tests/test_hostsim.py check EXACT == PRUNED == oracle ray for ray.
#pragma once
#include "dev_scene.h"

namespace wrt {

struct RayIn {
    float ox, oy, oz, dx, dy, dz, tmin, tmax;
};

struct VisitCounters {
    unsigned inner, leaf, tri, sph;
};

WRT_HD float sel3(int axis, float x, float y, float z)
{
    return axis == 0 ? x : (axis == 1 ? y : z);
}

// One axis of AABB::hit (AABB.cpp:16-30).  Returns false when the running interval became empty.
WRT_HD bool slab_axis(float l, float r, float o, float d, float& tmin, float& tmax)
{
    float inv = 1.f / d;
    float tn = (l - o) * inv;
    float tf = (r - o) * inv;
    if (tn > tf) { float s = tn; tn = tf; tf = s; }
    tmin = (tmin < tn) ? tn : tmin;   // std::max(tmin, tNear)
    tmax = (tf < tmax) ? tf : tmax;   // std::min(tmax, tFar)
    return !(tmin > tmax);
}

WRT_HD bool aabb_hit(const float lo[3], const float hi[3], const RayIn& r, float& t1, float& t2)
{
    float tmin = -WRT_INF, tmax = WRT_INF;
    if (!slab_axis(lo[0], hi[0], r.ox, r.dx, tmin, tmax)) return false;
    if (!slab_axis(lo[1], hi[1], r.oy, r.dy, tmin, tmax)) return false;
    if (!slab_axis(lo[2], hi[2], r.oz, r.dz, tmin, tmax)) return false;
    t1 = tmin; t2 = tmax;
    return true;
}

// Triangle::hit up to the acceptance of t (triangle.cpp:24-76).  e1 = p0-p1, e2 = p0-p2.
WRT_HD bool triangle_t(float p0x, float p0y, float p0z, float A, float B, float C,
                                           float D, float E, float F, const RayIn& r, float& t_out)
{
    const float G = r.dx, H = r.dy, I = r.dz;
    const float J = p0x - r.ox, K = p0y - r.oy, L = p0z - r.oz;
    const float EIHF = E * I - H * F;
    const float GFDI = G * F - D * I;
    const float DHEG = D * H - E * G;
    const float denom = A * EIHF + B * GFDI + C * DHEG;
    const float beta = (J * EIHF + K * GFDI + L * DHEG) / denom;
    if (beta < -WRT_EPS || beta > 1.f) return false;
    const float AKJB = A * K - J * B;
    const float JCAL = J * C - A * L;
    const float BLKC = B * L - K * C;
    const float gamma = (I * AKJB + H * JCAL + G * BLKC) / denom;
    if (gamma < -WRT_EPS || beta + gamma > 1.f) return false;
    const float t = -(F * AKJB + E * JCAL + D * BLKC) / denom;
    if (!(t > WRT_EPS)) return false;          // cmp(t) <= 0
    if (t < r.tmin || t > r.tmax) return false;
    t_out = t;
    return true;
}

// Sphere::hit up to the acceptance of t (sphere.cpp:19-68).
WRT_HD bool sphere_t(float cx, float cy, float cz, float radius, const float lo[3],
                                         const float hi[3], const RayIn& r, float& t_out, int& inside_out)
{
    float b1, b2;
    if (!aabb_hit(lo, hi, r, b1, b2)) return false;
    const float ocx = cx - r.ox, ocy = cy - r.oy, ocz = cz - r.oz;
    const float l_oc = ocx * ocx + ocy * ocy + ocz * ocz;
    const bool inside_flag = sqrtf(l_oc) < radius + WRT_EPS;
    const float t_ca = ocx * r.dx + ocy * r.dy + ocz * r.dz;
    if (t_ca < -WRT_EPS && !inside_flag) return false;
    const float t_hc = radius * radius - l_oc + t_ca * t_ca;
    if (!(t_hc > WRT_EPS)) return false;       // cmp(t_hc) <= 0
    const float d = sqrtf(t_hc);
    const float t1 = t_ca - d, t2 = t_ca + d;
    if (!(t2 > WRT_EPS)) return false;         // cmp(t2) <= 0
    float t; int inside;
    if (!(t1 > WRT_EPS)) { t = t2; inside = 1; } else { t = t1; inside = 0; }
    if (t < r.tmin || t > r.tmax) return false;
    t_out = t; inside_out = inside;
    return true;
}

#define WRT_PRUNE_REL 1.0001f

// ---- traversal state machine ------------------------------------------------------------------------
// One KDtreeAccel::traverse call, cut into the steps the kernels schedule: begin (root box, invDir),
// visit a node (optional PRUNED skip), interior step (:325-358), leaf (:359-374), pop (:375-384).
// kd_traverse() below runs them in a plain loop; the persistent kernels (trace_persistent.cuh) run the
// SAME steps but let idle lanes of a warp pick up new rays in between.  Either way every ray performs
// the same arithmetic in the same order.
// A ray for which a kd interval can be NaN or infinite: zero / tiny / infinite / NaN direction components (invDir infinite,
// huge, zero or NaN) or a non-finite origin.  Axis-parallel rays are the common case.  PRUNED traversal does not prune
// for such rays (box_prunable): with t = (split - o) * (+-inf) the reference's intervals become NaN / inf, its stack is
// no longer ordered by tmin, and the rule "stop the WHOLE traversal at the first popped entry with ray.tmax < tmin"
// (KDtreeAccel.cpp:323) can fire on an entry that lives INSIDE a sub-tree — skipping that sub-tree would let the traversal
// run on and find hits the reference never sees (found by the grazing-ray generator, tests/engines.py).
WRT_HD bool ray_is_degenerate(const RayIn& r, float ix, float iy, float iz)
{
    const float big = 1.0e30f;     // |invDir| below this: (split - o) * invDir cannot overflow for coordinates < 1e7 (INF)
    const bool inv_ok = fabsf(ix) < big && fabsf(iy) < big && fabsf(iz) < big && ix != 0.f && iy != 0.f && iz != 0.f;
    const bool org_ok = fabsf(r.ox) < big && fabsf(r.oy) < big && fabsf(r.oz) < big;
    return !(inv_ok && org_ok);
}

struct Trav {
    float tmin, tmax, best;
    float ix, iy, iz;     // invDir
    int node, sp, res;
    bool degen;           // ray_is_degenerate(): keep the NaN-guarded bounds test
};

struct TravStack {            // KDTodo (KDtreeAccel.h:47-51): one 16-byte entry = one local-memory access
    float4 e[WRT_STACK_DEPTH];    // node (int bits), tmin, tmax, unused
};

WRT_HD bool trav_begin(const DevSceneView& sc, const RayIn& r, Trav& T)
{
    if (!aabb_hit(sc.root_lo, sc.root_hi, r, T.tmin, T.tmax)) return false;     // :311-313
    T.ix = 1.f / r.dx; T.iy = 1.f / r.dy; T.iz = 1.f / r.dz;                     // invDir, :315
    T.node = 0; T.sp = 0; T.res = -1; T.best = WRT_INF;
    T.degen = ray_is_degenerate(r, T.ix, T.iy, T.iz);
    return true;
}

// Entry and exit distance of a "regular" ray (finite origin, finite non-zero invDir: ray_is_degenerate() == false) through a
// conservative box.  No slab can be NaN then ((finite or +-inf bound - finite) * finite non-zero), so no NaN guards are needed.
WRT_HD void box_interval_regular(float lox, float loy, float loz, float hix, float hiy, float hiz, const RayIn& r,
                                 float ix, float iy, float iz, float& entry, float& exit_)
{
    const float x0 = (lox - r.ox) * ix, x1 = (hix - r.ox) * ix;
    const float y0 = (loy - r.oy) * iy, y1 = (hiy - r.oy) * iy;
    const float z0 = (loz - r.oz) * iz, z1 = (hiz - r.oz) * iz;
    entry = fmaxf(fminf(x0, x1), fmaxf(fminf(y0, y1), fminf(z0, z1)));
    exit_ = fminf(fmaxf(x0, x1), fminf(fmaxf(y0, y1), fmaxf(z0, z1)));
}

// The PRUNED skip test on an explicit conservative box: nothing inside can change the traversal state when the
// ray (a) enters the box later than best * (1 + 1e-4) or (b) does not enter it at all / only behind its origin.
WRT_HD bool box_prunable(float lox, float loy, float loz, float hix, float hiy, float hiz, const RayIn& r,
                         float ix, float iy, float iz, int res, float best, bool degenerate)
{
    if (degenerate) return false;         // see ray_is_degenerate(): such rays take the reference's full traversal
    float en, ex;
    box_interval_regular(lox, loy, loz, hix, hiy, hiz, r, ix, iy, iz, en, ex);
    if (res >= 0 && en > best * WRT_PRUNE_REL) return true;
    const float m = 1e-4f * (fabsf(en) + fabsf(ex)) + 1e-4f;
    return (en > ex + m) || (ex < -m);
}

// The PRUNED skip test of a node: its conservative bounds (second half of the 32-byte node) are prunable.
// (a) alone — skipping only what lies behind the best hit — gave 345 Mrays/s on C3; (b), which also works before any
// hit is known, 1104 (profiles/r1_experiments.md, "Miss pruning").
template <bool PRUNED>
WRT_HD bool trav_skip(const DevSceneView& sc, const float4 na, const RayIn& r, const Trav& T)
{
    if (PRUNED) {
        const float4 nb = ldg4(&sc.nodes[2 * T.node + 1]);
        return box_prunable(na.z, na.w, nb.x, nb.y, nb.z, nb.w, r, T.ix, T.iy, T.iz, T.res, T.best, T.degen);
    }
    return false;
}

WRT_HD void trav_interior(const float4 na, const RayIn& r, Trav& T, TravStack& S)
{
    const unsigned packed = f2u(na.y);
    const int axis = (int)(packed & 3u);
    const float split = na.x;
    const float o_a = sel3(axis, r.ox, r.oy, r.oz);
    const float d_a = sel3(axis, r.dx, r.dy, r.dz);
    const float i_a = sel3(axis, T.ix, T.iy, T.iz);
    const float t = (split - o_a) * i_a;                                          // :328
    const bool below_first = (o_a < split) || (o_a == split && d_a <= 0.f);       // :331-332
    const int pair = (int)(packed >> 2);
    const int near_n = pair + (below_first ? 0 : 1);
    const int far_n = pair + (below_first ? 1 : 0);
    if (t > T.tmax || t <= 0.f) T.node = near_n;                                  // :345-346
    else if (t < T.tmin) T.node = far_n;                                          // :347-348
    else {                                                                        // :349-357
        if (T.sp < WRT_STACK_DEPTH) { S.e[T.sp] = make_float4(i2f(far_n), t, T.tmax, 0.f); ++T.sp; }
        T.node = near_n;
        T.tmax = t;
    }
}

// Leaf records (:359-374).  Besides one record per referenced primitive, in the reference's list order, a leaf holds
// "skip" records (kind 2, scene_layout.cpp): the conservative box of the next `n` records.  PRUNED traversal jumps
// over those records when the box is prunable — same argument as for sub-trees, and the order of the records that
// ARE tested is unchanged; EXACT traversal ignores skip records.
template <bool PRUNED, bool COUNT>
WRT_HD void trav_leaf(const DevSceneView& sc, const float4 na, const RayIn& r, Trav& T, VisitCounters* vc)
{
    const int first = f2i(na.x);
    const int cnt = (int)(f2u(na.y) >> 2);
    for (int i = 0; i < cnt; i++) {
        const float4* rec = sc.leaf_recs + 3 * (size_t)(first + i);
        const float4 r0 = ldg4(rec), r1 = ldg4(rec + 1), r2 = ldg4(rec + 2);
        float t; bool hit;
        if (f2i(r2.w) == WRT_REC_SKIP) {
            if (PRUNED) {
                if (COUNT) vc->inner++;      // a bounds test, counted with the node visits
                if (box_prunable(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r, T.ix, T.iy, T.iz, T.res, T.best, T.degen)) i += f2i(r0.w);
            }
            continue;
        }
        if (f2i(r2.w) == 0) {
            if (COUNT) vc->tri++;
            hit = triangle_t(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z, r, t);
        } else {
            if (COUNT) vc->sph++;
            const float lo[3] = { r1.y, r1.z, r1.w }, hi[3] = { r2.x, r2.y, r2.z };
            int inside;
            hit = sphere_t(r0.x, r0.y, r0.z, r1.x, lo, hi, r, t, inside);
        }
        if (hit && (t - T.best < -WRT_EPS)) { T.best = t; T.res = f2i(r0.w); }    // :367-371
    }
}

WRT_HD bool trav_pop(Trav& T, const TravStack& S)
{
    if (T.sp <= 0) return false;
    --T.sp;
    const float4 q = S.e[T.sp];
    T.node = f2i(q.x); T.tmin = q.y; T.tmax = q.z;
    return true;
}

// Returns the winning primitive id (index in Scene::objs) or -1, and its t in best_t.
template <bool PRUNED, bool COUNT>
WRT_HD int kd_traverse(const DevSceneView& sc, const RayIn& r, float& best_t, VisitCounters* vc)
{
    Trav T;
    TravStack S;
    if (!trav_begin(sc, r, T)) { best_t = WRT_INF; return -1; }
    for (;;) {
        if (r.tmax < T.tmin) break;                                               // :323
        const float4 na = ldg4(&sc.nodes[2 * T.node]);
        if (!trav_skip<PRUNED>(sc, na, r, T)) {
            if ((f2u(na.y) & 3u) != WRT_LEAF_TAG) {
                if (COUNT) vc->inner++;
                trav_interior(na, r, T, S);
                continue;
            }
            if (COUNT) vc->leaf++;
            trav_leaf<PRUNED, COUNT>(sc, na, r, T, vc);
        }
        if (!trav_pop(T, S)) break;
    }
    best_t = (T.res >= 0) ? T.best : WRT_INF;
    return T.res;
}

// The rest of Intersection for the winner: Scene::intersect re-runs g->hit (scene.cpp:26-27).
struct HitInfo {
    float px, py, pz, nx, ny, nz;
    int inside, matid;
};

WRT_HD void fill_hit(const DevSceneView& sc, int prim, const RayIn& r, float t, HitInfo& h)
{
    const float4 q0 = ldg4(&sc.prims[3 * (size_t)prim]);
    const float4 q1 = ldg4(&sc.prims[3 * (size_t)prim + 1]);
    const float4 q2 = ldg4(&sc.prims[3 * (size_t)prim + 2]);
    h.matid = f2i(q0.w);
    h.px = r.ox + r.dx * t; h.py = r.oy + r.dy * t; h.pz = r.oz + r.dz * t;      // ray(t)
    if (f2i(q1.w) == WRT_PRIM_TRIANGLE) {
        // n = (p1-p0) x (p2-p0), normalised, not flipped (triangle.cpp:79-84)
        const float ax = q1.x - q0.x, ay = q1.y - q0.y, az = q1.z - q0.z;
        const float bx = q2.x - q0.x, by = q2.y - q0.y, bz = q2.z - q0.z;
        float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        const float len = sqrtf(nx * nx + ny * ny + nz * nz);
        nx /= len; ny /= len; nz /= len;
        h.nx = nx; h.ny = ny; h.nz = nz;
        h.inside = ((r.dx * nx + r.dy * ny + r.dz * nz) < WRT_EPS) ? 0 : 1;
    } else {
        // sphere.cpp:52-76: inside flag from the chosen root, n = normalize(p - c)
        const float ocx = q0.x - r.ox, ocy = q0.y - r.oy, ocz = q0.z - r.oz;
        const float t_ca = ocx * r.dx + ocy * r.dy + ocz * r.dz;
        const float l_oc = ocx * ocx + ocy * ocy + ocz * ocz;
        const float t_hc = q1.x * q1.x - l_oc + t_ca * t_ca;
        const float t1 = t_ca - sqrtf(t_hc);
        h.inside = !(t1 > WRT_EPS) ? 1 : 0;
        float nx = h.px - q0.x, ny = h.py - q0.y, nz = h.pz - q0.z;
        const float len = sqrtf(nx * nx + ny * ny + nz * nz);
        h.nx = nx / len; h.ny = ny / len; h.nz = nz / len;
    }
}

// Ray(origin, dir): Vector3::normalize divides each component by sqrt(x*x+y*y+z*z) (vector.h:62-66).
WRT_HD void make_ray(float ox, float oy, float oz, float dx, float dy, float dz, RayIn& r)
{
    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
    r.ox = ox; r.oy = oy; r.oz = oz;
    r.dx = dx / len; r.dy = dy / len; r.dz = dz / len;
    r.tmin = 0.f; r.tmax = WRT_INF;
}

// Scene::shadowRayTest (scene.cpp:55-69): visible iff nothing is hit or the hit point equals p
// component-wise within EPS (Vector3 ==, vector.cpp:41-45).
template <bool PRUNED>
WRT_HD bool shadow_visible(const DevSceneView& sc, const RayIn& r, float px, float py, float pz)
{
    float t;
    const int prim = kd_traverse<PRUNED, false>(sc, r, t, nullptr);
    if (prim < 0) return true;
    const float hx = r.ox + r.dx * t, hy = r.oy + r.dy * t, hz = r.oz + r.dz * t;
    const float ex = hx - px, ey = hy - py, ez = hz - pz;
    const bool eq = !(ex < -WRT_EPS) && !(ex > WRT_EPS) && !(ey < -WRT_EPS) && !(ey > WRT_EPS) &&
                    !(ez < -WRT_EPS) && !(ez > WRT_EPS);
    return eq;
}

}  // namespace wrt
