// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
//
// C-ABI harness over the UNMODIFIED reference renderer (compiled where it lies under
// /root/reference by oracle/Makefile; outputs go to oracle/_ref/).  It lets the tests and the
// CPU-baseline leg of bench.py drive the reference's own classes:
//
//   Scene::intersect / Scene::occluded        R/src/scene/scene.cpp:21-81
//   KDtreeAccel::init / buildTree / traverse  R/src/scene/KDtreeAccel.cpp:12-388
//   Triangle::hit / Sphere::hit / AABB::hit   R/src/geometry/{triangle,sphere,AABB}.cpp
//   PathIntegrator / BidirPathTracing render  R/src/surfaceIntegrator/*.cpp
//
// (R = /root/reference/Winmad-s-raytracer-v1.0).  All reference members used here are public.
// The traversal call counter is installed with `ld --wrap`, so no reference source is edited.
#include "scene/scene.h"
#include "surfaceIntegrator/pathIntegrator.h"
#include "surfaceIntegrator/bidirPathTracing.h"
#include "surfaceIntegrator/whitted.h"
#include "material/bsdf.h"
#include "material/fresnel.h"
#include "sampler/sampler.h"
#include <unordered_map>
#include <vector>
#include <cstdint>
#include <cstring>

// ---- ld --wrap counter around KDtreeAccel::traverse(const Ray&, KDtreeAccelNode*) ----------
static unsigned long long g_traverse_calls = 0;
extern "C" Geometry* __real__ZN11KDtreeAccel8traverseERK3RayP15KDtreeAccelNode(
    KDtreeAccel*, const Ray&, KDtreeAccelNode*);
extern "C" Geometry* __wrap__ZN11KDtreeAccel8traverseERK3RayP15KDtreeAccelNode(
    KDtreeAccel* self, const Ray& ray, KDtreeAccelNode* tr)
{
    ++g_traverse_calls;
    return __real__ZN11KDtreeAccel8traverseERK3RayP15KDtreeAccelNode(self, ray, tr);
}

namespace {

struct RefHandle {
    int kind;  // 0 = path tracer, 1 = bidirectional path tracer, 2 = Whitted
    PathIntegrator* pt;
    BidirPathTracing* bpt;
    WhittedIntegrator* wh;
    std::unordered_map<const Geometry*, int> index_of;
    SurfaceIntegrator* integ() { return kind == 0 ? (SurfaceIntegrator*)pt : kind == 1 ? (SurfaceIntegrator*)bpt : (SurfaceIntegrator*)wh; }
    Scene& scene() { return integ()->scene; }
};

void finish_scene(RefHandle* h, int width, int height)
{
    Scene& sc = h->scene();
    // same steps as Scene::init after loadScene (R/src/scene/scene.cpp:476-488)
    if (sc.objs.size() > 0) {
        sc.kdtreeAccel.init(sc.objs);
        sc.kdtreeAccel.buildTree(sc.kdtreeAccel.root, 1);
        Vector3 diag = sc.kdtreeAccel.root->box.r - sc.kdtreeAccel.root->box.l;
        Real diameter2 = diag.sqrLength();
        sc.sceneSphere.sceneCenter = (sc.kdtreeAccel.root->box.l + sc.kdtreeAccel.root->box.r) * 0.5f;
        sc.sceneSphere.sceneRadius = std::sqrt(diameter2) * 0.5f;
        sc.sceneSphere.invSceneRadiusSqr = 1.f / diameter2;
    }
    h->index_of.clear();
    for (size_t i = 0; i < sc.objs.size(); i++) h->index_of[sc.objs[i]] = (int)i;
    SurfaceIntegrator* in = h->integ();
    in->width = width; in->height = height;
    in->film = new ImageFilm(height, width);
    if (h->kind == 1) h->bpt->pixelNum = width * height;
}

void count_tree(KDtreeAccelNode* tr, int dep, long long* nnodes, long long* nrefs, int* depth)
{
    if (!tr) return;
    ++*nnodes;
    if (dep > *depth) *depth = dep;
    if (tr->axis == -1) { *nrefs += tr->objNum; return; }
    count_tree(tr->left, dep + 1, nnodes, nrefs, depth);
    count_tree(tr->right, dep + 1, nnodes, nrefs, depth);
}

struct FlatOut {
    int* axis; float* split; int* left; int* right; int* first_ref; int* nref; float* box6; int* refs;
    int nnodes; long long nrefs;
};

int flatten(RefHandle* h, KDtreeAccelNode* tr, FlatOut& o)
{
    int me = o.nnodes++;
    o.axis[me] = tr->axis; o.split[me] = tr->splitPlane;
    o.box6[6 * me + 0] = tr->box.l.x; o.box6[6 * me + 1] = tr->box.l.y; o.box6[6 * me + 2] = tr->box.l.z;
    o.box6[6 * me + 3] = tr->box.r.x; o.box6[6 * me + 4] = tr->box.r.y; o.box6[6 * me + 5] = tr->box.r.z;
    if (tr->axis == -1) {
        o.left[me] = o.right[me] = -1;
        o.first_ref[me] = (int)o.nrefs; o.nref[me] = tr->objNum;
        for (int i = 0; i < tr->objNum; i++) o.refs[o.nrefs++] = h->index_of[tr->objlist[i]];
    } else {
        o.first_ref[me] = -1; o.nref[me] = tr->objNum;
        int l = flatten(h, tr->left, o);
        int r = flatten(h, tr->right, o);
        o.left[me] = l; o.right[me] = r;
    }
    return me;
}

inline Ray make_ray(const float* r8)
{
    Ray r;  // default ctor: no normalisation; fields are then exactly "a Ray after its ctor"
    r.origin = Vector3(r8[0], r8[1], r8[2]);
    r.dir = Vector3(r8[3], r8[4], r8[5]);
    r.tmin = r8[6]; r.tmax = r8[7];
    return r;
}

}  // namespace

extern "C" {

void* ref_create(int kind)
{
    RefHandle* h = new RefHandle();
    h->kind = kind; h->pt = 0; h->bpt = 0; h->wh = 0;
    if (kind == 2) {
        h->wh = new WhittedIntegrator();
        h->wh->maxTracingDepth = 7; h->wh->samplesPerPixel = 1;
    } else if (kind == 0) {
        h->pt = new PathIntegrator();
        h->pt->maxTracingDepth = 7; h->pt->samplesPerPixel = 1;
        h->pt->samplesOfLight = 8; h->pt->samplesOfHemisphere = 4;
    } else {
        h->bpt = new BidirPathTracing();
        h->bpt->minPathLength = 0; h->bpt->maxPathLength = 10; h->bpt->iterations = 1;
        h->bpt->samplesPerPixel = 1; h->bpt->controlLength = 3;
    }
    return h;
}

// Reference scene loader (TinyXML + tinyobjloader) exactly as Scene::init does it.
int ref_load_scene_file(void* hv, const char* path, int width, int height)
{
    RefHandle* h = (RefHandle*)hv;
    std::vector<char> buf(path, path + strlen(path) + 1);
    h->scene().loadScene(&buf[0]);
    finish_scene(h, width, height);
    return (int)h->scene().objs.size();
}

// Programmatic scene: same object/material/light construction calls as the loader makes.
// materials: nm x 11 floats (diffuse rgb, phong rgb, phongExp, specular rgb, index)
// prims: np x (kind, 9 floats, matid); kind 0 = triangle p0 p1 p2, kind 1 = sphere c.xyz r
// lights: nl x 12 floats (p0 p1 p2 intensity rgb)
// cam: pos3 forward3 up3 xres yres fov
int ref_build_scene(void* hv, int nm, const float* materials, int np, const int* kind,
                    const float* data9, const int* matid, int nl, const float* lights12,
                    const float* cam12, int width, int height)
{
    RefHandle* h = (RefHandle*)hv;
    Scene& sc = h->scene();
    for (int i = 0; i < nm; i++) {
        const float* m = materials + 11 * i;
        Material mat;
        mat.diffuse = Color3(m[0], m[1], m[2]); mat.phong = Color3(m[3], m[4], m[5]);
        mat.phongExp = m[6]; mat.specular = Color3(m[7], m[8], m[9]); mat.index = m[10];
        sc.addMaterial(mat);
    }
    for (int i = 0; i < np; i++) {
        const float* d = data9 + 9 * i;
        if (kind[i] == 0)
            sc.addGeometry(new Triangle(Vector3(d[0], d[1], d[2]), Vector3(d[3], d[4], d[5]),
                                        Vector3(d[6], d[7], d[8]), matid[i]));
        else
            sc.addGeometry(new Sphere(Vector3(d[0], d[1], d[2]), d[3], matid[i]));
    }
    for (int i = 0; i < nl; i++) {
        const float* l = lights12 + 12 * i;
        sc.addLight(new AreaLight(Vector3(l[0], l[1], l[2]), Vector3(l[3], l[4], l[5]),
                                  Vector3(l[6], l[7], l[8]), Color3(l[9], l[10], l[11])));
    }
    sc.camera.setup(Vector3(cam12[0], cam12[1], cam12[2]), Vector3(cam12[3], cam12[4], cam12[5]),
                    Vector3(cam12[6], cam12[7], cam12[8]), cam12[9], cam12[10], cam12[11]);
    finish_scene(h, width, height);
    return np;
}

int ref_num_prims(void* hv) { return (int)((RefHandle*)hv)->scene().objs.size(); }
int ref_num_materials(void* hv) { return (int)((RefHandle*)hv)->scene().materials.size(); }
int ref_num_lights(void* hv) { return (int)((RefHandle*)hv)->scene().lights.size(); }

void ref_get_prims(void* hv, int* kind, float* data9, int* matid)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    for (size_t i = 0; i < sc.objs.size(); i++) {
        float* d = data9 + 9 * i;
        for (int k = 0; k < 9; k++) d[k] = 0.f;
        if (Triangle* t = dynamic_cast<Triangle*>(sc.objs[i])) {
            kind[i] = 0;
            d[0] = t->p0.x; d[1] = t->p0.y; d[2] = t->p0.z;
            d[3] = t->p1.x; d[4] = t->p1.y; d[5] = t->p1.z;
            d[6] = t->p2.x; d[7] = t->p2.y; d[8] = t->p2.z;
        } else if (Sphere* s = dynamic_cast<Sphere*>(sc.objs[i])) {
            kind[i] = 1;
            d[0] = s->center.x; d[1] = s->center.y; d[2] = s->center.z; d[3] = s->radius;
        } else kind[i] = -1;
        matid[i] = sc.objs[i]->getMatId();
    }
}

void ref_get_prim_boxes(void* hv, float* box6)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    for (size_t i = 0; i < sc.objs.size(); i++) {
        const AABB& b = sc.objs[i]->box;
        box6[6 * i + 0] = b.l.x; box6[6 * i + 1] = b.l.y; box6[6 * i + 2] = b.l.z;
        box6[6 * i + 3] = b.r.x; box6[6 * i + 4] = b.r.y; box6[6 * i + 5] = b.r.z;
    }
}

void ref_get_materials(void* hv, float* m11)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    for (size_t i = 0; i < sc.materials.size(); i++) {
        const Material& mat = sc.materials[i];
        float* m = m11 + 11 * i;
        m[0] = mat.diffuse.r; m[1] = mat.diffuse.g; m[2] = mat.diffuse.b;
        m[3] = mat.phong.r; m[4] = mat.phong.g; m[5] = mat.phong.b; m[6] = mat.phongExp;
        m[7] = mat.specular.r; m[8] = mat.specular.g; m[9] = mat.specular.b; m[10] = mat.index;
    }
}

// per light: p0 d1 d2 intensity invArea frame.x frame.y frame.z  = 3+3+3+3+1+9 = 22 floats
void ref_get_lights(void* hv, float* l22)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    for (size_t i = 0; i < sc.lights.size(); i++) {
        AreaLight* a = dynamic_cast<AreaLight*>(sc.lights[i]);
        float* l = l22 + 22 * i;
        if (!a) { for (int k = 0; k < 22; k++) l[k] = 0; continue; }
        l[0] = a->p0.x; l[1] = a->p0.y; l[2] = a->p0.z;
        l[3] = a->d1.x; l[4] = a->d1.y; l[5] = a->d1.z;
        l[6] = a->d2.x; l[7] = a->d2.y; l[8] = a->d2.z;
        l[9] = a->intensity.r; l[10] = a->intensity.g; l[11] = a->intensity.b;
        l[12] = a->invArea;
        l[13] = a->localFrame.x.x; l[14] = a->localFrame.x.y; l[15] = a->localFrame.x.z;
        l[16] = a->localFrame.y.x; l[17] = a->localFrame.y.y; l[18] = a->localFrame.y.z;
        l[19] = a->localFrame.z.x; l[20] = a->localFrame.z.y; l[21] = a->localFrame.z.z;
    }
}

// pos3 forward3 up3 xres yres fov imagePlaneDist rasterToWorld[16] worldToRaster[16] = 45 floats
void ref_get_camera(void* hv, float* c45)
{
    Camera& c = ((RefHandle*)hv)->scene().camera;
    c45[0] = c.pos.x; c45[1] = c.pos.y; c45[2] = c.pos.z;
    c45[3] = c.forward.x; c45[4] = c.forward.y; c45[5] = c.forward.z;
    c45[6] = c.up.x; c45[7] = c.up.y; c45[8] = c.up.z;
    c45[9] = c.xResolution; c45[10] = c.yResolution; c45[11] = c.horizontalFOV;
    c45[12] = c.imagePlaneDist;
    memcpy(c45 + 13, c.rasterToWorld.m.m, 16 * sizeof(float));
    memcpy(c45 + 29, c.worldToRaster.m.m, 16 * sizeof(float));
}

void ref_get_scene_sphere(void* hv, float* s5)
{
    SceneSphere& s = ((RefHandle*)hv)->scene().sceneSphere;
    s5[0] = s.sceneCenter.x; s5[1] = s.sceneCenter.y; s5[2] = s.sceneCenter.z;
    s5[3] = s.sceneRadius; s5[4] = s.invSceneRadiusSqr;
}

void ref_tree_stats(void* hv, long long* nnodes, long long* nrefs, int* depth, int* depmax)
{
    RefHandle* h = (RefHandle*)hv;
    *nnodes = 0; *nrefs = 0; *depth = 0;
    count_tree(h->scene().kdtreeAccel.root, 1, nnodes, nrefs, depth);
    *depmax = h->scene().kdtreeAccel.depMax;
}

// DFS pre-order flatten of the reference's pointer tree; leaf reference lists keep objlist order.
void ref_tree_flatten(void* hv, int* axis, float* split, int* left, int* right, int* first_ref,
                      int* nref, float* box6, int* refs)
{
    RefHandle* h = (RefHandle*)hv;
    FlatOut o = { axis, split, left, right, first_ref, nref, box6, refs, 0, 0 };
    flatten(h, h->scene().kdtreeAccel.root, o);
}

// Scene::intersect(ray, inter) on a batch of already-constructed rays (8 floats each).
// prim = index in scene.objs or -1.  Optional outputs may be NULL.
void ref_intersect(void* hv, const float* rays8, long long n, int* prim, float* t,
                   float* p3, float* n3, int* inside, int* matid)
{
    RefHandle* h = (RefHandle*)hv;
    Scene& sc = h->scene();
    for (long long i = 0; i < n; i++) {
        Ray r = make_ray(rays8 + 8 * i);
        Intersection inter;
        inter.t = INF; inter.inside = 0; inter.matId = 0;
        Geometry* g = sc.intersect(r, inter);
        prim[i] = g ? h->index_of[g] : -1;
        if (t) t[i] = g ? inter.t : INF;
        if (g) {
            if (p3) { p3[3 * i] = inter.p.x; p3[3 * i + 1] = inter.p.y; p3[3 * i + 2] = inter.p.z; }
            if (n3) { n3[3 * i] = inter.n.x; n3[3 * i + 1] = inter.n.y; n3[3 * i + 2] = inter.n.z; }
            if (inside) inside[i] = inter.inside;
            if (matid) matid[i] = inter.matId;
        } else {
            if (p3) { p3[3 * i] = p3[3 * i + 1] = p3[3 * i + 2] = 0.f; }
            if (n3) { n3[3 * i] = n3[3 * i + 1] = n3[3 * i + 2] = 0.f; }
            if (inside) inside[i] = 0;
            if (matid) matid[i] = 0;
        }
    }
}

// Scene::occluded(p1, dir, p2): 9 floats per query (the Ray ctor inside normalises dir).
void ref_occluded(void* hv, const float* q9, long long n, unsigned char* occ)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    for (long long i = 0; i < n; i++) {
        const float* q = q9 + 9 * i;
        occ[i] = sc.occluded(Vector3(q[0], q[1], q[2]), Vector3(q[3], q[4], q[5]),
                             Vector3(q[6], q[7], q[8])) ? 1 : 0;
    }
}

// Ray(origin, dir) constructor (normalises dir): 6 floats in, 8 floats out.
void ref_make_rays(const float* od6, long long n, float* rays8)
{
    for (long long i = 0; i < n; i++) {
        const float* q = od6 + 6 * i;
        Ray r(Vector3(q[0], q[1], q[2]), Vector3(q[3], q[4], q[5]));
        float* o = rays8 + 8 * i;
        o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z;
        o[3] = r.dir.x; o[4] = r.dir.y; o[5] = r.dir.z; o[6] = r.tmin; o[7] = r.tmax;
    }
}

// Camera::generateRay(x, y) for a batch of raster positions.
void ref_generate_rays(void* hv, const float* xy, long long n, float* rays8)
{
    Camera& c = ((RefHandle*)hv)->scene().camera;
    for (long long i = 0; i < n; i++) {
        Ray r = c.generateRay(xy[2 * i], xy[2 * i + 1]);
        float* o = rays8 + 8 * i;
        o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z;
        o[3] = r.dir.x; o[4] = r.dir.y; o[5] = r.dir.z; o[6] = r.tmin; o[7] = r.tmax;
    }
}

// Primitive known-answer tests.
int ref_triangle_hit(const float* tri9, const float* ray8, float* t, float* p3, float* n3, int* inside)
{
    Triangle tr(Vector3(tri9[0], tri9[1], tri9[2]), Vector3(tri9[3], tri9[4], tri9[5]),
                Vector3(tri9[6], tri9[7], tri9[8]), 1);
    Ray r = make_ray(ray8);
    Intersection inter; inter.t = INF; inter.inside = 0;
    bool hit = tr.hit(r, inter);
    *t = inter.t;
    if (hit) {
        p3[0] = inter.p.x; p3[1] = inter.p.y; p3[2] = inter.p.z;
        n3[0] = inter.n.x; n3[1] = inter.n.y; n3[2] = inter.n.z; *inside = inter.inside;
    }
    return hit ? 1 : 0;
}

int ref_sphere_hit(const float* cr4, const float* ray8, float* t, float* p3, float* n3, int* inside)
{
    Sphere s(Vector3(cr4[0], cr4[1], cr4[2]), cr4[3], 1);
    Ray r = make_ray(ray8);
    Intersection inter; inter.t = INF; inter.inside = 0;
    bool hit = s.hit(r, inter);
    *t = inter.t;
    if (hit) {
        p3[0] = inter.p.x; p3[1] = inter.p.y; p3[2] = inter.p.z;
        n3[0] = inter.n.x; n3[1] = inter.n.y; n3[2] = inter.n.z; *inside = inter.inside;
    }
    return hit ? 1 : 0;
}

int ref_aabb_hit(const float* box6, const float* ray8, float* t1, float* t2)
{
    AABB b; b.l = Vector3(box6[0], box6[1], box6[2]); b.r = Vector3(box6[3], box6[4], box6[5]);
    Ray r = make_ray(ray8);
    Real a = 0, c = 0;
    bool hit = b.hit(r, a, c);
    *t1 = a; *t2 = c;
    return hit ? 1 : 0;
}

unsigned long long ref_traverse_calls(void) { return g_traverse_calls; }
void ref_reset_traverse_calls(void) { g_traverse_calls = 0; }

static void copy_film(ImageFilm* film, float* out)
{
    for (int i = 0; i < film->height; i++)
        for (int j = 0; j < film->width; j++) {
            const Color3& c = film->color[i][j];
            float* o = out + 3 * ((size_t)i * film->width + j);
            o[0] = c.r; o[1] = c.g; o[2] = c.b;
        }
}

static void clear_film(ImageFilm* film)
{
    for (int i = 0; i < film->height; i++)
        for (int j = 0; j < film->width; j++) film->color[i][j] = Color3(0.f);
}

// PathIntegrator: SurfaceIntegrator::render() (pixel loop, 1/spp scale).  film: H x W x 3 linear.
int ref_render_pt(void* hv, int spp, int max_depth, unsigned seed, float* film)
{
    RefHandle* h = (RefHandle*)hv;
    if (h->kind != 0) return -1;
    h->pt->samplesPerPixel = spp; h->pt->maxTracingDepth = max_depth;
    h->pt->rng.seed(seed); h->pt->rng.mti = RNG::N;  // force a regenerate from the new seed state
    clear_film(h->pt->film);
    h->pt->render();
    copy_film(h->pt->film, film);
    return 0;
}

// WhittedIntegrator: SurfaceIntegrator::render() with WhittedIntegrator::raytracing (R/src/surfaceIntegrator/whitted.cpp).
// The light pick uses libc rand() (:28), seeded here as well.
int ref_render_whitted(void* hv, int spp, int max_depth, unsigned seed, float* film)
{
    RefHandle* h = (RefHandle*)hv;
    if (h->kind != 2) return -1;
    h->wh->samplesPerPixel = spp; h->wh->maxTracingDepth = max_depth;
    h->wh->rng.seed(seed); h->wh->rng.mti = RNG::N;
    srand(seed);
    clear_film(h->wh->film);
    h->wh->render();
    copy_film(h->wh->film, film);
    return 0;
}

// PathIntegrator on a sub-window of rows [row0,row1) — for bounded CPU-baseline samples.  Same
// per-pixel body as SurfaceIntegrator::render (R/src/surfaceIntegrator/surfaceIntegrator.cpp:20-41).
int ref_render_pt_rows(void* hv, int spp, int max_depth, unsigned seed, int row0, int row1, int row_stride,
                       int col0, int col1, float* film)
{
    RefHandle* h = (RefHandle*)hv;
    if (h->kind != 0) return -1;
    PathIntegrator* in = h->pt;
    in->samplesPerPixel = spp; in->maxTracingDepth = max_depth;
    in->rng.seed(seed); in->rng.mti = RNG::N;
    if (film) clear_film(in->film);
    if (row_stride < 1) row_stride = 1;
    for (int i = row0; i < row1; i += row_stride)
        for (int j = col0; j < col1; j++)
            for (int k = 0; k < spp; k++) {
                Vector3 v0 = Vector3(j - 0.5f, i - 0.5f, 0);
                Vector3 v1 = Vector3(j + 0.5f, i - 0.5f, 0);
                Vector3 v2 = Vector3(j - 0.5f, i + 0.5f, 0);
                Vector3 posRaster = sampleRectangleStratified(in->rng.randVector3(), v0, v1, v2, k, spp);
                Ray ray = in->scene.camera.generateRay(posRaster.x, posRaster.y);
                Color3 tmp = in->raytracing(ray, 0);
                in->film->addColor(i, j, tmp);
            }
    if (film) { in->film->scale(1.f / spp); copy_film(in->film, film); }
    return 0;
}

// BidirPathTracing::render() with `iterations` runIteration calls.  The returned film is the raw
// accumulator film->color[a][b] (NOT transposed, NOT divided by iterations) — outputImage does both
// (R/src/surfaceIntegrator/bidirPathTracing.cpp:29-46).
int ref_render_bdpt(void* hv, int iterations, unsigned seed, int control_length, int max_path_length,
                    float* film)
{
    RefHandle* h = (RefHandle*)hv;
    if (h->kind != 1) return -1;
    h->bpt->iterations = iterations;
    h->bpt->controlLength = control_length;
    h->bpt->maxPathLength = max_path_length;
    h->bpt->rng.seed(seed); h->bpt->rng.mti = RNG::N;
    clear_film(h->bpt->film);
    h->bpt->render();
    copy_film(h->bpt->film, film);
    return 0;
}

// ---- shading known-answer helpers (BSDF / light / sampler), for the device shading tests ------
// in: wi3 (world, pointing away from the surface), n3, matid, wo3; out: f rgb, cosWo, dirPdf, revPdf,
// continueProb, isDelta, valid
void ref_bsdf_eval(void* hv, const float* wi3, const float* n3, int matid, const float* wo3, float* out9)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    Intersection inter; inter.t = 1.f; inter.p = Vector3(0.f); inter.n = Vector3(n3[0], n3[1], n3[2]);
    inter.inside = 0; inter.matId = matid;
    BSDF b(Vector3(wi3[0], wi3[1], wi3[2]), inter, sc);
    for (int k = 0; k < 9; k++) out9[k] = 0.f;
    out9[8] = b.isValid() ? 1.f : 0.f;
    if (!b.isValid() || matid <= 0) return;
    Real cosWo = 0, dp = 0, rp = 0;
    Color3 f = b.f(sc, Vector3(wo3[0], wo3[1], wo3[2]), cosWo, &dp, &rp);
    out9[0] = f.r; out9[1] = f.g; out9[2] = f.b; out9[3] = cosWo; out9[4] = dp; out9[5] = rp;
    out9[6] = b.continueProb; out9[7] = b.isDelta ? 1.f : 0.f;
}

// BSDF::sample: out = f rgb, wo3, pdf, cosWo, sampledType
void ref_bsdf_sample(void* hv, const float* wi3, const float* n3, int matid, const float* rand3, float* out9)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    Intersection inter; inter.t = 1.f; inter.p = Vector3(0.f); inter.n = Vector3(n3[0], n3[1], n3[2]);
    inter.inside = 0; inter.matId = matid;
    BSDF b(Vector3(wi3[0], wi3[1], wi3[2]), inter, sc);
    for (int k = 0; k < 9; k++) out9[k] = 0.f;
    if (!b.isValid() || matid <= 0) return;
    Vector3 wo(0.f); Real pdf = 0, cosWo = 0; int type = 0;
    Color3 f = b.sample(sc, Vector3(rand3[0], rand3[1], rand3[2]), wo, pdf, cosWo, &type);
    out9[0] = f.r; out9[1] = f.g; out9[2] = f.b; out9[3] = wo.x; out9[4] = wo.y; out9[5] = wo.z;
    out9[6] = pdf; out9[7] = cosWo; out9[8] = (float)type;
}

// AreaLight::illuminance: out = illu rgb, dirToLight3, dist, directPdf, emissionPdf, cosAtLight
void ref_light_illuminance(void* hv, int light, const float* pos3, const float* rand3, float* out10)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    Vector3 d(0.f); Real dist = 0, dpdf = 0, epdf = 0, cosl = 0;
    Color3 c = sc.lights[light]->illuminance(sc.sceneSphere, Vector3(pos3[0], pos3[1], pos3[2]),
        Vector3(rand3[0], rand3[1], rand3[2]), d, dist, dpdf, &epdf, &cosl);
    out10[0] = c.r; out10[1] = c.g; out10[2] = c.b; out10[3] = d.x; out10[4] = d.y; out10[5] = d.z;
    out10[6] = dist; out10[7] = dpdf; out10[8] = epdf; out10[9] = cosl;
}

// AreaLight::emit: out = energy rgb, pos3, dir3, emissionPdf, directPdfArea, cosAtLight
void ref_light_emit(void* hv, int light, const float* dirRand3, const float* posRand3, float* out12)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    Vector3 p(0.f), d(0.f); Real epdf = 0, dpdf = 0, cosl = 0;
    Color3 c = sc.lights[light]->emit(sc.sceneSphere, Vector3(dirRand3[0], dirRand3[1], dirRand3[2]),
        Vector3(posRand3[0], posRand3[1], posRand3[2]), p, d, epdf, &dpdf, &cosl);
    out12[0] = c.r; out12[1] = c.g; out12[2] = c.b; out12[3] = p.x; out12[4] = p.y; out12[5] = p.z;
    out12[6] = d.x; out12[7] = d.y; out12[8] = d.z; out12[9] = epdf; out12[10] = dpdf; out12[11] = cosl;
}


// ---- batched shading known-answer tests -----------------------------------------------------------------------
// One call evaluates n inputs of one reference function; the device (wrt_debug_shading) and hostsim take the same
// (what, in, out) layout, so the three can be compared entry by entry.  in/out strides are fixed per `what`:
//   0 BSDF::f            in 10: wi3 n3 matid wo3         out 9: f.rgb cosWo dirPdf revPdf continueProb isDelta valid
//   1 BSDF::sample       in 10: wi3 n3 matid rand3       out 9: f.rgb wo3 pdf cosWo type
//   2 BSDF::pdf          in 10: wi3 n3 matid wo3         out 2: pdf, reverse pdf
//   3 AreaLight::illuminance  in 7: light pos3 rand3     out 10: illu.rgb dir3 dist directPdf emissionPdf cosAtLight
//   4 AreaLight::emit    in 7: light dirRand3 posRand3   out 12: energy.rgb pos3 dir3 emissionPdf directPdfArea cosAtLight
//   5 AreaLight::getRadiance  in 4: light dir3           out 5: rad.rgb directPdfArea emissionPdf
//   6 fresnelDielectric  in 2: cosI index                out 1
//   7 samplers           in 13: rand3 power v1 v2 v3     out 13: sampleTriangle p3 | sampleCosHemisphere dir3 pdf |
//                                                                 samplePowerCosHemisphere dir3 | cosHemispherePdf(z, cos dir) |
//                                                                 powerCosHemispherePdf(z, power dir, power) | 0
//   8 camera sample      in 5: rand.x rand.y i j k       out 8: the Ray of SurfaceIntegrator::render's sample k of pixel (i, j)
//                        (sampleRectangleStratified + Camera::generateRay; spp = `iparam`)
int ref_shading_in_stride(int what) { static const int s[9] = { 10, 10, 10, 7, 7, 4, 2, 13, 5 }; return what >= 0 && what < 9 ? s[what] : 0; }
int ref_shading_out_stride(int what) { static const int s[9] = { 9, 9, 2, 10, 12, 5, 1, 13, 8 }; return what >= 0 && what < 9 ? s[what] : 0; }

int ref_shading_batch(void* hv, int what, int iparam, const float* in, long long n, float* out)
{
    RefHandle* h = (RefHandle*)hv;
    Scene& sc = h->scene();
    const int is = ref_shading_in_stride(what), os = ref_shading_out_stride(what);
    if (!is) return -1;
    for (long long e = 0; e < n; e++) {
        const float* a = in + (size_t)is * e;
        float* o = out + (size_t)os * e;
        for (int k = 0; k < os; k++) o[k] = 0.f;
        if (what <= 2) {
            Intersection inter; inter.t = 1.f; inter.p = Vector3(0.f); inter.n = Vector3(a[3], a[4], a[5]);
            inter.inside = 0; inter.matId = (int)a[6];
            BSDF b(Vector3(a[0], a[1], a[2]), inter, sc);
            if (what == 0) {
                o[8] = b.isValid() ? 1.f : 0.f;
                if (!b.isValid() || inter.matId <= 0) continue;
                Real cosWo = 0, dp = 0, rp = 0;
                Color3 f = b.f(sc, Vector3(a[7], a[8], a[9]), cosWo, &dp, &rp);
                o[0] = f.r; o[1] = f.g; o[2] = f.b; o[3] = cosWo; o[4] = dp; o[5] = rp;
                o[6] = b.continueProb; o[7] = b.isDelta ? 1.f : 0.f;
            } else if (what == 1) {
                if (!b.isValid() || inter.matId <= 0) continue;
                Vector3 wo(0.f); Real pdf = 0, cosWo = 0; int type = 0;
                Color3 f = b.sample(sc, Vector3(a[7], a[8], a[9]), wo, pdf, cosWo, &type);
                o[0] = f.r; o[1] = f.g; o[2] = f.b; o[3] = wo.x; o[4] = wo.y; o[5] = wo.z; o[6] = pdf; o[7] = cosWo; o[8] = (float)type;
            } else {
                if (!b.isValid() || inter.matId <= 0) continue;
                o[0] = b.pdf(sc, Vector3(a[7], a[8], a[9]), false);
                o[1] = b.pdf(sc, Vector3(a[7], a[8], a[9]), true);
            }
        } else if (what == 3) {
            Vector3 d(0.f); Real dist = 0, dpdf = 0, epdf = 0, cosl = 0;
            Color3 c = sc.lights[(int)a[0]]->illuminance(sc.sceneSphere, Vector3(a[1], a[2], a[3]), Vector3(a[4], a[5], a[6]), d, dist, dpdf, &epdf, &cosl);
            o[0] = c.r; o[1] = c.g; o[2] = c.b; o[3] = d.x; o[4] = d.y; o[5] = d.z; o[6] = dist; o[7] = dpdf; o[8] = epdf; o[9] = cosl;
        } else if (what == 4) {
            Vector3 p(0.f), d(0.f); Real epdf = 0, dpdf = 0, cosl = 0;
            Color3 c = sc.lights[(int)a[0]]->emit(sc.sceneSphere, Vector3(a[1], a[2], a[3]), Vector3(a[4], a[5], a[6]), p, d, epdf, &dpdf, &cosl);
            o[0] = c.r; o[1] = c.g; o[2] = c.b; o[3] = p.x; o[4] = p.y; o[5] = p.z; o[6] = d.x; o[7] = d.y; o[8] = d.z; o[9] = epdf; o[10] = dpdf; o[11] = cosl;
        } else if (what == 5) {
            Real dpa = 0, epdf = 0;
            Color3 c = sc.lights[(int)a[0]]->getRadiance(sc.sceneSphere, Vector3(a[1], a[2], a[3]), Vector3(0.f), &dpa, &epdf);
            o[0] = c.r; o[1] = c.g; o[2] = c.b; o[3] = dpa; o[4] = epdf;
        } else if (what == 6) {
            o[0] = fresnelDielectric(a[0], a[1]);
        } else if (what == 7) {
            const Vector3 r(a[0], a[1], a[2]);
            Vector3 p = sampleTriangle(r, Vector3(a[4], a[5], a[6]), Vector3(a[7], a[8], a[9]), Vector3(a[10], a[11], a[12]));
            o[0] = p.x; o[1] = p.y; o[2] = p.z;
            Real pdf = 0;
            Vector3 c = sampleCosHemisphere(r, &pdf);
            o[3] = c.x; o[4] = c.y; o[5] = c.z; o[6] = pdf;
            Vector3 g = samplePowerCosHemisphere(r, a[3], NULL);
            o[7] = g.x; o[8] = g.y; o[9] = g.z;
            o[10] = cosHemispherePdf(Vector3(0.f, 0.f, 1.f), c);
            o[11] = powerCosHemispherePdf(Vector3(0.f, 0.f, 1.f), g, a[3]);
        } else if (what == 8) {
            const int i = (int)a[2], j = (int)a[3], k = (int)a[4];
            Vector3 v0 = Vector3(j - 0.5f, i - 0.5f, 0), v1 = Vector3(j + 0.5f, i - 0.5f, 0), v2 = Vector3(j - 0.5f, i + 0.5f, 0);
            Vector3 pr = sampleRectangleStratified(Vector3(a[0], a[1], 0.f), v0, v1, v2, k, iparam);
            Ray r = sc.camera.generateRay(pr.x, pr.y);
            o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z; o[3] = r.dir.x; o[4] = r.dir.y; o[5] = r.dir.z; o[6] = r.tmin; o[7] = r.tmax;
        }
    }
    return 0;
}

// Scene::shadowRayTest(ray, p) (scene.cpp:55-69) and bool Scene::intersect(ray) (scene.cpp:45-53) on batches of
// already-constructed rays.
void ref_shadow_test(void* hv, const float* rays8, const float* target3, long long n, float* visible)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    for (long long i = 0; i < n; i++) {
        Ray r = make_ray(rays8 + 8 * i);
        visible[i] = sc.shadowRayTest(r, Vector3(target3[3 * i], target3[3 * i + 1], target3[3 * i + 2]));
    }
}

void ref_intersect_any(void* hv, const float* rays8, long long n, unsigned char* hit)
{
    Scene& sc = ((RefHandle*)hv)->scene();
    for (long long i = 0; i < n; i++) { Ray r = make_ray(rays8 + 8 * i); hit[i] = sc.intersect(r) ? 1 : 0; }
}

// ---- random-number tapes: same random numbers for the reference and the device ------------------------------------
// The reference draws every number of a render from ONE MT19937 stream, sample after sample.  These entry points run the
// unmodified per-sample code and record, for every sample, the `stride` floats the stream holds from the point where
// the sample starts (a sample consumes a prefix of them).  The device integrators replay the tape (wrt_debug_set_rng_tape),
// so both sides follow the SAME path sample for sample and the films can be compared per pixel, without Monte-Carlo noise.
static void tape_record(const RNG& at, int stride, float* dst)
{
    RNG copy = at;                      // 624 words + index: the stream from here on
    for (int k = 0; k < stride; k++) dst[k] = copy.randFloat();
}
static inline int stream_pos(const RNG& r) { return r.mti % RNG::N; }

// PathIntegrator: SurfaceIntegrator::render()'s loop (surfaceIntegrator.cpp:17-45) with the recorder around each sample.
// tape: [H*W*spp][stride]; sample_rgb: [H*W*spp][3] (raytracing()'s return value per sample); draws: [H*W*spp].
int ref_render_pt_tape(void* hv, int spp, int max_depth, unsigned seed, int stride, float* tape, float* sample_rgb,
                       int* draws, float* film)
{
    RefHandle* h = (RefHandle*)hv;
    if (h->kind != 0) return -1;
    PathIntegrator* in = h->pt;
    in->samplesPerPixel = spp; in->maxTracingDepth = max_depth;
    in->rng.seed(seed); in->rng.mti = RNG::N;
    clear_film(in->film);
    size_t s = 0;
    for (int i = 0; i < in->height; i++)
        for (int j = 0; j < in->width; j++)
            for (int k = 0; k < spp; k++, s++) {
                tape_record(in->rng, stride, tape + s * (size_t)stride);
                const int p0 = stream_pos(in->rng);
                Vector3 v0 = Vector3(j - 0.5f, i - 0.5f, 0);
                Vector3 v1 = Vector3(j + 0.5f, i - 0.5f, 0);
                Vector3 v2 = Vector3(j - 0.5f, i + 0.5f, 0);
                Vector3 posRaster = sampleRectangleStratified(in->rng.randVector3(), v0, v1, v2, k, spp);
                Ray ray = in->scene.camera.generateRay(posRaster.x, posRaster.y);
                Color3 tmp = in->raytracing(ray, 0);
                in->film->addColor(i, j, tmp);
                if (sample_rgb) { sample_rgb[3 * s] = tmp.r; sample_rgb[3 * s + 1] = tmp.g; sample_rgb[3 * s + 2] = tmp.b; }
                if (draws) draws[s] = (stream_pos(in->rng) - p0 + RNG::N) % RNG::N;
            }
    in->film->scale(1.f / spp);
    if (film) copy_film(in->film, film);
    return 0;
}

// MT19937 stream of the reference RNG (R/src/math/rng.cpp): n floats after seeding.
void ref_rng_floats(unsigned seed, int n, float* out)
{
    RNG rng(seed);
    rng.mti = RNG::N;
    for (int i = 0; i < n; i++) out[i] = rng.randFloat();
}

}  // extern "C"
