/* TEST INFRASTRUCTURE ONLY — plain-C restatement ("port") of the reference's hot path, used as the
 * portable checker when oracle/_ref (the compiled reference itself) is not available.
 *
 * Parity status: PINNED — this port is checked bit-for-bit (primitive ids, t, occlusion flags)
 * against the compiled reference on every scene/batch in tests/golden/ (tests/test_oracle.py) and the
 * committed golden vectors were produced by the reference itself (tests/golden/make_golden.py).
 *
 * Restates, function by function (R = /root/reference/Winmad-s-raytracer-v1.0/src):
 *   cmp                      R/math/math.cpp:8-11          EPS = 1e-3f, INF = 1e7f (math.h:17-18)
 *   Vector3::normalize       R/math/vector.h:62-66         Ray ctor R/geometry/ray.h:14-16
 *   AABB::hit                R/geometry/AABB.cpp:9-32
 *   Triangle::hit            R/geometry/triangle.cpp:22-87
 *   Sphere::hit              R/geometry/sphere.cpp:17-78
 *   KDtreeAccel::traverse    R/scene/KDtreeAccel.cpp:309-388 (no early exit; t - best < -EPS rule)
 *   Scene::intersect         R/scene/scene.cpp:21-43       (re-runs hit() on the winner)
 *   Scene::shadowRayTest / occluded   R/scene/scene.cpp:55-81
 * Compile with -ffp-contract=off and without -march=native: float results must be the no-FMA ones.
 */
#include <math.h>
#include <stddef.h>
#include "port.h"

#define EPSF 1e-3f
#define INFF 1e7f

typedef struct { float o[3], d[3], tmin, tmax; } ray_t;
typedef struct { float t, p[3], n[3]; int inside, matid; } inter_t;

static int cmpf(float x) { return (x < -EPSF) ? -1 : (x > EPSF); }

static void normalize3(float v[3])
{
    float len = (float)sqrt((double)(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    v[0] /= len; v[1] /= len; v[2] /= len;
}

static int aabb_hit(const float l[3], const float r[3], const ray_t* ray, float* t1, float* t2)
{
    float tmin = -INFF, tmax = INFF;
    int i;
    for (i = 0; i < 3; i++) {
        float inv = 1.f / ray->d[i];
        float tn = (l[i] - ray->o[i]) * inv;
        float tf = (r[i] - ray->o[i]) * inv;
        if (tn > tf) { float s = tn; tn = tf; tf = s; }
        tmin = (tmin < tn) ? tn : tmin;      /* std::max(tmin, tNear) */
        tmax = (tf < tmax) ? tf : tmax;      /* std::min(tmax, tFar)  */
        if (tmin > tmax) return 0;
    }
    *t1 = tmin; *t2 = tmax;
    return 1;
}

static int triangle_hit(const float* q, int matid, const ray_t* ray, inter_t* in)
{
    const float* p0 = q; const float* p1 = q + 3; const float* p2 = q + 6;
    float A = p0[0] - p1[0], B = p0[1] - p1[1], C = p0[2] - p1[2];
    float D = p0[0] - p2[0], E = p0[1] - p2[1], F = p0[2] - p2[2];
    float G = ray->d[0], H = ray->d[1], I = ray->d[2];
    float J = p0[0] - ray->o[0], K = p0[1] - ray->o[1], L = p0[2] - ray->o[2];
    float EIHF = E * I - H * F, GFDI = G * F - D * I, DHEG = D * H - E * G;
    float denom = (A * EIHF + B * GFDI + C * DHEG);
    float beta = (J * EIHF + K * GFDI + L * DHEG) / denom;
    float AKJB, JCAL, BLKC, gamma, a[3], b[3];
    if (cmpf(beta) < 0 || beta > 1.f) { in->t = INFF; return 0; }
    AKJB = A * K - J * B; JCAL = J * C - A * L; BLKC = B * L - K * C;
    gamma = (I * AKJB + H * JCAL + G * BLKC) / denom;
    if (cmpf(gamma) < 0 || beta + gamma > 1.f) { in->t = INFF; return 0; }
    in->t = -(F * AKJB + E * JCAL + D * BLKC) / denom;
    if (cmpf(in->t) <= 0) { in->t = INFF; return 0; }
    if (in->t < ray->tmin || in->t > ray->tmax) { in->t = INFF; return 0; }
    in->p[0] = ray->o[0] + ray->d[0] * in->t; in->p[1] = ray->o[1] + ray->d[1] * in->t; in->p[2] = ray->o[2] + ray->d[2] * in->t;
    a[0] = p1[0] - p0[0]; a[1] = p1[1] - p0[1]; a[2] = p1[2] - p0[2];
    b[0] = p2[0] - p0[0]; b[1] = p2[1] - p0[1]; b[2] = p2[2] - p0[2];
    in->n[0] = a[1] * b[2] - a[2] * b[1]; in->n[1] = a[2] * b[0] - a[0] * b[2]; in->n[2] = a[0] * b[1] - a[1] * b[0];
    normalize3(in->n);
    in->inside = ((ray->d[0] * in->n[0] + ray->d[1] * in->n[1] + ray->d[2] * in->n[2]) < EPSF) ? 0 : 1;
    in->matid = matid;
    return 1;
}

static void sphere_box(const float* q, float l[3], float r[3])  /* Sphere::setBox + AABB::extend */
{
    int a;
    for (a = 0; a < 3; a++) { l[a] = q[a] - q[3]; r[a] = q[a] + q[3]; }
    for (a = 0; a < 3; a++) if (cmpf(l[a] - r[a]) == 0) r[a] += 10 * EPSF;
}

static int sphere_hit(const float* q, int matid, const ray_t* ray, inter_t* in)
{
    float l[3], r[3], b1, b2, oc[3], l_oc, t_ca, t_hc, t1, t2, d;
    int inside_flag = 0;
    sphere_box(q, l, r);
    if (!aabb_hit(l, r, ray, &b1, &b2)) { in->t = INFF; return 0; }
    oc[0] = q[0] - ray->o[0]; oc[1] = q[1] - ray->o[1]; oc[2] = q[2] - ray->o[2];
    l_oc = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2];
    if ((float)sqrt((double)l_oc) < q[3] + EPSF) inside_flag = 1;
    t_ca = oc[0] * ray->d[0] + oc[1] * ray->d[1] + oc[2] * ray->d[2];
    if (cmpf(t_ca) < 0 && !inside_flag) { in->t = INFF; return 0; }
    t_hc = q[3] * q[3] - l_oc + t_ca * t_ca;
    if (cmpf(t_hc) <= 0) { in->t = INFF; return 0; }
    d = (float)sqrt((double)t_hc);
    t1 = t_ca - d; t2 = t_ca + d;
    if (cmpf(t2) <= 0) { in->t = INFF; return 0; }
    if (cmpf(t1) <= 0) { in->t = t2; in->inside = 1; } else { in->t = t1; in->inside = 0; }
    if (in->t < ray->tmin || in->t > ray->tmax) { in->t = INFF; return 0; }
    in->p[0] = ray->o[0] + ray->d[0] * in->t; in->p[1] = ray->o[1] + ray->d[1] * in->t; in->p[2] = ray->o[2] + ray->d[2] * in->t;
    in->n[0] = in->p[0] - q[0]; in->n[1] = in->p[1] - q[1]; in->n[2] = in->p[2] - q[2];
    normalize3(in->n);
    in->matid = matid;
    return 1;
}

static int prim_hit(const port_scene* s, int p, const ray_t* ray, inter_t* in, unsigned long long* c)
{
    if (s->kind[p] == 0) { if (c) c[2]++; return triangle_hit(s->data9 + 9 * (size_t)p, s->matid[p], ray, in); }
    if (c) c[3]++;
    return sphere_hit(s->data9 + 9 * (size_t)p, s->matid[p], ray, in);
}

/* KDtreeAccel::traverse */
static int traverse(const port_scene* s, const ray_t* ray, unsigned long long* c)
{
    float tmin, tmax, inv[3], tmp = INFF;
    int todo_node[64]; float todo_tmin[64], todo_tmax[64];
    int todo_pos = 0, res = -1, tr = 0;
    if (!aabb_hit(s->root_box, s->root_box + 3, ray, &tmin, &tmax)) return -1;
    inv[0] = 1.f / ray->d[0]; inv[1] = 1.f / ray->d[1]; inv[2] = 1.f / ray->d[2];
    while (tr >= 0) {
        if (ray->tmax < tmin) break;
        if (s->axis[tr] != -1) {
            int axis = s->axis[tr], near_n, far_n, below_first;
            float t = (s->split[tr] - ray->o[axis]) * inv[axis];
            if (c) c[0]++;
            below_first = (ray->o[axis] < s->split[tr]) || (ray->o[axis] == s->split[tr] && ray->d[axis] <= 0);
            if (below_first) { near_n = s->left[tr]; far_n = s->right[tr]; }
            else { near_n = s->right[tr]; far_n = s->left[tr]; }
            if (t > tmax || t <= 0) tr = near_n;
            else if (t < tmin) tr = far_n;
            else {
                todo_node[todo_pos] = far_n; todo_tmin[todo_pos] = t; todo_tmax[todo_pos] = tmax; ++todo_pos;
                tr = near_n; tmax = t;
            }
        } else {
            inter_t in;
            int i;
            if (c) c[1]++;
            in.t = INFF;
            for (i = 0; i < s->n_ref[tr]; i++) {
                int p = s->refs[s->first_ref[tr] + i];
                if (prim_hit(s, p, ray, &in, c)) {
                    if (cmpf(in.t - tmp) < 0) { tmp = in.t; res = p; }
                }
            }
            if (todo_pos > 0) { --todo_pos; tr = todo_node[todo_pos]; tmin = todo_tmin[todo_pos]; tmax = todo_tmax[todo_pos]; }
            else break;
        }
    }
    return res;
}

static void load_ray(const float* r8, ray_t* r)
{
    r->o[0] = r8[0]; r->o[1] = r8[1]; r->o[2] = r8[2]; r->d[0] = r8[3]; r->d[1] = r8[4]; r->d[2] = r8[5];
    r->tmin = r8[6]; r->tmax = r8[7];
}

void port_intersect(const port_scene* s, const float* rays8, long long n, int32_t* prim, float* t,
                    float* p3, float* n3, int32_t* inside, int32_t* matid, unsigned long long* counters)
{
    long long i;
    for (i = 0; i < n; i++) {
        ray_t ray; inter_t in;
        int g;
        load_ray(rays8 + 8 * i, &ray);
        in.t = INFF; in.inside = 0; in.matid = 0;
        in.p[0] = in.p[1] = in.p[2] = in.n[0] = in.n[1] = in.n[2] = 0.f;
        g = traverse(s, &ray, counters);
        if (g >= 0) prim_hit(s, g, &ray, &in, NULL);      /* Scene::intersect re-runs hit() */
        prim[i] = g;
        if (t) t[i] = (g >= 0) ? in.t : INFF;
        if (p3) { p3[3 * i] = in.p[0]; p3[3 * i + 1] = in.p[1]; p3[3 * i + 2] = in.p[2]; }
        if (n3) { n3[3 * i] = in.n[0]; n3[3 * i + 1] = in.n[1]; n3[3 * i + 2] = in.n[2]; }
        if (inside) inside[i] = (g >= 0) ? in.inside : 0;
        if (matid) matid[i] = (g >= 0) ? in.matid : 0;
    }
}

void port_make_rays(const float* q, long long n, float* r8)
{
    long long i;
    for (i = 0; i < n; i++) {
        float d[3];
        d[0] = q[6 * i + 3]; d[1] = q[6 * i + 4]; d[2] = q[6 * i + 5];
        normalize3(d);
        r8[8 * i] = q[6 * i]; r8[8 * i + 1] = q[6 * i + 1]; r8[8 * i + 2] = q[6 * i + 2];
        r8[8 * i + 3] = d[0]; r8[8 * i + 4] = d[1]; r8[8 * i + 5] = d[2]; r8[8 * i + 6] = 0.f; r8[8 * i + 7] = INFF;
    }
}

void port_occluded(const port_scene* s, const float* q9, long long n, uint8_t* occluded)
{
    long long i;
    for (i = 0; i < n; i++) {
        const float* q = q9 + 9 * i;
        ray_t ray; inter_t in;
        int g, eq;
        float hp[3];
        ray.o[0] = q[0]; ray.o[1] = q[1]; ray.o[2] = q[2]; ray.d[0] = q[3]; ray.d[1] = q[4]; ray.d[2] = q[5];
        normalize3(ray.d); ray.tmin = 0.f; ray.tmax = INFF;
        g = traverse(s, &ray, NULL);
        if (g < 0) { occluded[i] = 0; continue; }
        prim_hit(s, g, &ray, &in, NULL);
        hp[0] = ray.o[0] + ray.d[0] * in.t; hp[1] = ray.o[1] + ray.d[1] * in.t; hp[2] = ray.o[2] + ray.d[2] * in.t;
        eq = cmpf(hp[0] - q[6]) == 0 && cmpf(hp[1] - q[7]) == 0 && cmpf(hp[2] - q[8]) == 0;
        occluded[i] = eq ? 0 : 1;
    }
}

/* Scene::shadowRayTest(ray, p), scene.cpp:55-69: the ray is taken as given (no re-normalisation); 1.0 = visible. */
void port_shadow_test(const port_scene* s, const float* rays8, const float* target3, long long n, float* visible)
{
    long long i;
    for (i = 0; i < n; i++) {
        ray_t ray; inter_t in;
        int g, eq;
        float hp[3];
        const float* p = target3 + 3 * i;
        load_ray(rays8 + 8 * i, &ray);
        g = traverse(s, &ray, NULL);
        if (g < 0) { visible[i] = 1.0f; continue; }
        prim_hit(s, g, &ray, &in, NULL);
        hp[0] = ray.o[0] + ray.d[0] * in.t; hp[1] = ray.o[1] + ray.d[1] * in.t; hp[2] = ray.o[2] + ray.d[2] * in.t;
        eq = cmpf(hp[0] - p[0]) == 0 && cmpf(hp[1] - p[1]) == 0 && cmpf(hp[2] - p[2]) == 0;
        visible[i] = eq ? 1.0f : 0.0f;
    }
}

/* bool Scene::intersect(ray), scene.cpp:45-53: a full closest-hit traversal, only hit / no hit is returned. */
void port_intersect_any(const port_scene* s, const float* rays8, long long n, uint8_t* hit)
{
    long long i;
    for (i = 0; i < n; i++) {
        ray_t ray;
        load_ray(rays8 + 8 * i, &ray);
        hit[i] = traverse(s, &ray, NULL) >= 0 ? 1 : 0;
    }
}

int port_triangle_hit(const float* tri9, const float* ray8, float* t)
{
    ray_t ray; inter_t in; int h;
    load_ray(ray8, &ray); in.t = INFF;
    h = triangle_hit(tri9, 1, &ray, &in);
    *t = in.t;
    return h;
}

int port_sphere_hit(const float* cr4, const float* ray8, float* t, int* inside)
{
    ray_t ray; inter_t in; int h;
    load_ray(ray8, &ray); in.t = INFF; in.inside = 0;
    h = sphere_hit(cr4, 1, &ray, &in);
    *t = in.t; *inside = in.inside;
    return h;
}

int port_aabb_hit(const float* box6, const float* ray8, float* t1, float* t2)
{
    ray_t ray;
    load_ray(ray8, &ray);
    *t1 = 0.f; *t2 = 0.f;
    return aabb_hit(box6, box6 + 3, &ray, t1, t2);
}
