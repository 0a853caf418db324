// Scene layout builder (pure host code, no CUDA calls): turns the reference's KD-tree (as arrays,
// include/wrt.h wrt_kdtree) and Scene::objs / materials / lights into the HBM layout of dev_scene.h.
//
//  * nodes are re-laid out breadth-first with sibling pairs adjacent (one child index per node,
//    top levels first in memory);
//  * every leaf's objlist becomes a contiguous run of 48-byte leaf records (triangle edges p0-p1,
//    p0-p2 pre-subtracted exactly as Triangle::hit does, triangle.cpp:24-30);
//  * every node gets a conservative box of all primitives referenced in its sub-tree, used only by
//    the PRUNED traversal (traverse.cuh);
//  * leaves with more than a handful of primitives get "skip" records: the list is cut into chunks of 3-5
//    consecutive entries (and, for long lists, groups of 4 chunks), each preceded by a record holding the
//    conservative box of the chunk, so PRUNED traversal can jump over chunks the ray cannot hit without
//    changing the order of the tests it does perform.  (The reference tree stops splitting early — the median
//    reference of the 1 M-triangle scene sits in a leaf of 14 primitives, 24 when spheres are mixed in.)
// wrt_scene_create (scene_upload.cu) uploads the result; the test-only hostsim library points a
// DevSceneView at the host vectors instead.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>
#include "scene_layout.h"

namespace wrt {

namespace {

// The per-primitive and per-leaf phases of build_layout write disjoint ranges: they run on threads (WRT_LAYOUT_THREADS, default = the
// machine's, at most 16; 1 = serial) and produce the same bytes (tests/test_host.py::test_parallel_layout_equals_serial).
int layout_threads(size_t work)
{
    int t = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("WRT_LAYOUT_THREADS")) t = atoi(e);
    if (t > 16) t = 16;
    if (t < 1 || work < 50000) t = 1;
    return t;
}

template <class F>
void parallel_ranges(int n, int threads, F fn)      // fn(i0, i1) over [0, n) cut into `threads` ranges
{
    if (threads <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back([=, &fn] { fn((int)((long long)n * t / threads), (int)((long long)n * (t + 1) / threads)); });
    fn(0, (int)((long long)n / threads));
    for (auto& x : th) x.join();
}

inline float as_float(int32_t v) { float f; memcpy(&f, &v, 4); return f; }
inline float as_float_u(uint32_t v) { float f; memcpy(&f, &v, 4); return f; }

inline float round_down(double v) { float f = (float)v; if ((double)f > v) f = nextafterf(f, -INFINITY); return f; }
inline float round_up(double v) { float f = (float)v; if ((double)f < v) f = nextafterf(f, INFINITY); return f; }

// Conservative box of every point a reference hit on this primitive can report.
// Triangle: accepted hits have beta, gamma >= -EPS and beta+gamma <= 1 (triangle.cpp:47,58), i.e.
// they lie in the triangle blown up in barycentric space; we use twice that slack plus a relative pad.
void conservative_box(int kind, const float* d, double lo[3], double hi[3])
{
    if (kind == WRT_PRIM_TRIANGLE) {
        const double e = 2e-3;
        const double bc[3][2] = { { -e, -e }, { 1 + 3 * e, -e }, { -e, 1 + 3 * e } };
        for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; }
        for (int c = 0; c < 3; c++)
            for (int a = 0; a < 3; a++) {
                double p0 = d[a], p1 = d[3 + a], p2 = d[6 + a];
                double v = p0 + bc[c][0] * (p1 - p0) + bc[c][1] * (p2 - p0);
                lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v);
            }
    } else {
        const double r = std::fabs((double)d[3]) * (1 + 2e-3) + 2e-3;
        for (int a = 0; a < 3; a++) { lo[a] = d[a] - r; hi[a] = d[a] + r; }
    }
    for (int a = 0; a < 3; a++) {
        double pad = 1e-5 * std::max(std::fabs(lo[a]), std::fabs(hi[a])) + 1e-6;
        lo[a] -= pad; hi[a] += pad;
    }
}

// Skip-record plan of a leaf with c primitives: chunk sizes (empty = no skip records) and chunks per group.
constexpr int kSkipGroupMinChunks = 7; // leaves with at least this many chunks get a second level (groups of 4 chunks)

int env_int(const char* name, int dflt, int lo, int hi)
{
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    const int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}
// Tuning knobs, read once per build_layout call: WRT_LEAF_SKIP=0 disables skip records, WRT_LEAF_SKIP_MIN = smallest
// leaf that gets them (smaller leaves are tested record by record), WRT_LEAF_SKIP_CHUNK = target chunk length.
// Defaults measured on C3 / C5 (profiles/r1_experiments.md, "Skip records").
struct SkipCfg { bool enabled; int min_leaf, chunk; };

SkipCfg skip_config()
{
    const char* e = getenv("WRT_LEAF_SKIP");
    SkipCfg c;
    c.enabled = !(e && e[0] == '0');
    c.min_leaf = env_int("WRT_LEAF_SKIP_MIN", 16, 2, 1 << 20);
    c.chunk = env_int("WRT_LEAF_SKIP_CHUNK", 6, 2, 64);
    return c;
}

void plan_leaf(int c, const SkipCfg& cfg, std::vector<int>& chunks, int& per_group)
{
    chunks.clear(); per_group = 0;
    if (!cfg.enabled || c < cfg.min_leaf) return;
    const int g = cfg.chunk;
    const int k = std::max(1, (c + g / 2) / g);
    for (int j = 0; j < k; j++) chunks.push_back(c / k + (j < c % k ? 1 : 0));
    if (k >= kSkipGroupMinChunks) per_group = 4;
}

int64_t leaf_record_count(int c, const SkipCfg& cfg)
{
    std::vector<int> chunks; int per_group;
    plan_leaf(c, cfg, chunks, per_group);
    int64_t n = c + (int64_t)chunks.size();
    if (per_group) for (size_t j = 0; j < chunks.size(); j += per_group) if (chunks.size() - j >= 2) n++;
    return n;
}

void normalize3(float v[3])
{
    float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= len; v[1] /= len; v[2] /= len;
}
void cross3(const float a[3], const float b[3], float r[3])
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

// AreaLight ctor (light.h:90-104) + Frame::buildFromZ (frame.cpp:3-11)
void make_light(const float* l12, DevLight& L)
{
    for (int a = 0; a < 3; a++) { L.p0[a] = l12[a]; L.d1[a] = l12[3 + a] - l12[a]; L.d2[a] = l12[6 + a] - l12[a]; }
    float n[3]; cross3(L.d1, L.d2, n);
    float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    L.inv_area = 2.f / len;
    normalize3(n);
    float z[3] = { n[0], n[1], n[2] };
    normalize3(z);
    float tmpx[3] = { 1.f, 0.f, 0.f };
    if (std::fabs(z[0]) > 0.99f) { tmpx[0] = 0.f; tmpx[1] = 1.f; }
    float y[3]; cross3(z, tmpx, y); normalize3(y);
    float x[3]; cross3(y, z, x);
    for (int a = 0; a < 3; a++) { L.fx[a] = x[a]; L.fy[a] = y[a]; L.fz[a] = z[a]; L.intensity[a] = l12[9 + a]; }
}

}  // namespace

#define set_error(msg) do { err = (msg); } while (0)

bool build_layout(const wrt_scene_desc* d, SceneLayout& L, std::string& err)
{
    if (!d) { err = "null scene description"; return false; }
    const wrt_kdtree& T = d->tree;
    if (d->n_prims <= 0 || !d->prim_kind || !d->prim_data || !d->prim_matid) {
        set_error("scene has no primitives"); return false;
    }
    if (T.n_nodes <= 0 || !T.axis || !T.split || !T.left || !T.right || !T.first_ref || !T.n_ref ||
        (T.n_refs > 0 && !T.refs)) {
        set_error("kd-tree missing (build it on the host first)"); return false;
    }
    const int nn = T.n_nodes;

    // ---- breadth-first relayout with adjacent sibling pairs ------------------------------------
    std::vector<int32_t> order;            // new index -> old index
    std::vector<int32_t> pair_of(nn, -1);  // old interior index -> new index of its left child
    std::vector<char> seen(nn, 0);
    std::vector<int32_t> above(nn, 0);     // old index -> interior nodes on the way from the root (= stack entries a ray can hold there)
    int max_above = 0;
    order.reserve(nn);
    order.push_back(0); seen[0] = 1;
    for (size_t head = 0; head < order.size(); head++) {
        const int o = order[head];
        max_above = std::max(max_above, (int)above[o]);
        if (T.axis[o] == -1) continue;
        if (T.axis[o] < 0 || T.axis[o] > 2) { set_error("node axis out of range"); return false; }
        const int l = T.left[o], r = T.right[o];
        if (l < 0 || l >= nn || r < 0 || r >= nn || seen[l] || seen[r] || l == r) {
            set_error("malformed kd-tree (child index out of range or shared)");
            return false;
        }
        seen[l] = seen[r] = 1;
        above[l] = above[r] = above[o] + 1;
        pair_of[o] = (int32_t)order.size();
        order.push_back(l); order.push_back(r);
    }
    const int n_live = (int)order.size();
    // A ray holds at most one pending far child per interior node above it.  The traversal stacks (per-thread and the
    // pooled scheduler's scratch) have WRT_STACK_DEPTH entries; a deeper tree is refused here instead of silently
    // dropping far children in the kernels (the reference's own bound is depMax + 5 <= 29 for 1e8 primitives).
    if (max_above > WRT_STACK_DEPTH) {
        char buf[160];
        snprintf(buf, sizeof buf, "kd-tree too deep: %d interior levels, the traversal stack holds %d", max_above, WRT_STACK_DEPTH);
        set_error(buf); return false;
    }

    // ---- leaf records --------------------------------------------------------------------------
    const SkipCfg skips = skip_config();
    std::vector<int64_t> first_rec(n_live, 0);
    std::vector<int32_t> leaf_recs_n(n_live, 0);
    int64_t n_recs = 0;
    for (int i = 0; i < n_live; i++) {
        const int o = order[i];
        if (T.axis[o] != -1) continue;
        const int64_t f = T.first_ref[o], c = T.n_ref[o];
        if (c < 0 || (c > 0 && (f < 0 || f + c > T.n_refs))) { set_error("leaf reference range out of bounds"); return false; }
        const int64_t nr = leaf_record_count((int)c, skips);
        first_rec[i] = n_recs; leaf_recs_n[i] = (int32_t)nr; n_recs += nr;
    }
    if (n_recs >= (int64_t)1 << 31 || nn >= (1 << 29)) { set_error("tree too large"); return false; }

    std::vector<float> pboxes((size_t)d->n_prims * 6);  // reference prim boxes (spheres need theirs in hit())
    std::vector<float> cons((size_t)d->n_prims * 6);    // conservative boxes, rounded outward
    const int nthr = layout_threads((size_t)d->n_prims + (size_t)n_recs);
    std::atomic<int> bad(0);            // 1 unknown kind, 2 material id, 3 light id, 4 primitive reference, 5 record plan
    parallel_ranges(d->n_prims, nthr, [&](int i0, int i1) {
        for (int i = i0; i < i1; i++) {
            const float* pd = d->prim_data + 9 * (size_t)i;
            if (d->prim_kind[i] != WRT_PRIM_TRIANGLE && d->prim_kind[i] != WRT_PRIM_SPHERE) { bad = 1; return; }
            // the shaders index materials[matid] / lights[-matid-1] directly (the reference reads out of bounds here)
            const int m = d->prim_matid[i];
            if (m > 0 && m >= d->n_materials) { bad = 2; return; }
            if (m < 0 && (-(int64_t)m - 1) >= d->n_lights) { bad = 3; return; }
            float* b = &pboxes[6 * (size_t)i];
            if (d->prim_kind[i] == WRT_PRIM_SPHERE) {  // Sphere::setBox + AABB::extend
                for (int a = 0; a < 3; a++) { b[a] = pd[a] - pd[3]; b[3 + a] = pd[a] + pd[3]; }
                for (int a = 0; a < 3; a++) { float df = b[a] - b[3 + a]; if (!(df < -WRT_EPS) && !(df > WRT_EPS)) b[3 + a] += 10 * WRT_EPS; }
            }
            double lo[3], hi[3];
            conservative_box(d->prim_kind[i], pd, lo, hi);
            for (int a = 0; a < 3; a++) { cons[6 * (size_t)i + a] = round_down(lo[a]); cons[6 * (size_t)i + 3 + a] = round_up(hi[a]); }
        }
    });
    if (bad == 1) { set_error("unknown primitive kind"); return false; }
    if (bad == 2) { set_error("primitive material id past the material table"); return false; }
    if (bad == 3) { set_error("emitter primitive refers to a light past the light table"); return false; }

    std::vector<float4> recs((size_t)n_recs * 3);
    std::vector<float> nb((size_t)n_live * 6);  // node bounds
    for (int i = 0; i < n_live; i++) for (int a = 0; a < 3; a++) { nb[6 * (size_t)i + a] = INFINITY; nb[6 * (size_t)i + 3 + a] = -INFINITY; }
    // every leaf writes its own run of records and its own bounds: node ranges on threads
    parallel_ranges(n_live, nthr, [&](int n0, int n1) {
    std::vector<int> chunks;
    for (int i = n0; i < n1; i++) {
        const int o = order[i];
        if (T.axis[o] != -1) continue;
        const int c = T.n_ref[o];
        for (int k = 0; k < c; k++) {
            const int p = T.refs[T.first_ref[o] + k];
            if (p < 0 || p >= d->n_prims) { bad = 4; return; }
        }
        chunks.clear(); int per_group;
        plan_leaf(c, skips, chunks, per_group);
        int64_t pos = first_rec[i];
        auto emit_prim = [&](int k) {
            const int p = T.refs[T.first_ref[o] + k];
            const float* pd = d->prim_data + 9 * (size_t)p;
            float4* r = &recs[3 * (size_t)pos++];
            if (d->prim_kind[p] == WRT_PRIM_TRIANGLE) {
                r[0] = make_float4(pd[0], pd[1], pd[2], as_float(p));
                r[1] = make_float4(pd[0] - pd[3], pd[1] - pd[4], pd[2] - pd[5], 0.f);
                r[2] = make_float4(pd[0] - pd[6], pd[1] - pd[7], pd[2] - pd[8], as_float(0));
            } else {
                const float* b = &pboxes[6 * (size_t)p];
                r[0] = make_float4(pd[0], pd[1], pd[2], as_float(p));
                r[1] = make_float4(pd[3], b[0], b[1], b[2]);
                r[2] = make_float4(b[3], b[4], b[5], as_float(1));
            }
            for (int a = 0; a < 3; a++) {
                nb[6 * (size_t)i + a] = std::min(nb[6 * (size_t)i + a], cons[6 * (size_t)p + a]);
                nb[6 * (size_t)i + 3 + a] = std::max(nb[6 * (size_t)i + 3 + a], cons[6 * (size_t)p + 3 + a]);
            }
        };
        auto emit_skip = [&](int k0, int k1, int n_inside) {   // conservative box of list entries [k0, k1)
            float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
            for (int k = k0; k < k1; k++) {
                const int p = T.refs[T.first_ref[o] + k];
                for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], cons[6 * (size_t)p + a]); hi[a] = std::max(hi[a], cons[6 * (size_t)p + 3 + a]); }
            }
            float4* r = &recs[3 * (size_t)pos++];
            r[0] = make_float4(lo[0], lo[1], lo[2], as_float(n_inside));
            r[1] = make_float4(hi[0], hi[1], hi[2], 0.f);
            r[2] = make_float4(0.f, 0.f, 0.f, as_float(WRT_REC_SKIP));
        };
        if (chunks.empty()) for (int k = 0; k < c; k++) emit_prim(k);
        else {
            int k = 0;
            for (size_t j = 0; j < chunks.size();) {
                const size_t in_group = per_group ? std::min<size_t>(per_group, chunks.size() - j) : chunks.size() - j;
                if (per_group && in_group >= 2) {
                    int prims_in = 0;
                    for (size_t q = j; q < j + in_group; q++) prims_in += chunks[q];
                    emit_skip(k, k + prims_in, prims_in + (int)in_group);
                }
                for (size_t q = j; q < j + in_group; q++) {
                    emit_skip(k, k + chunks[q], chunks[q]);
                    for (int e = 0; e < chunks[q]; e++) emit_prim(k++);
                }
                j += in_group;
            }
        }
        if (pos != first_rec[i] + leaf_recs_n[i]) { bad = 5; return; }
    }
    });
    if (bad == 4) { set_error("leaf references a primitive out of range"); return false; }
    if (bad == 5) { set_error("internal: leaf record plan mismatch"); return false; }
    for (int i = n_live - 1; i >= 0; i--) {  // children have larger indices than parents
        const int o = order[i];
        if (T.axis[o] == -1) continue;
        for (int c = 0; c < 2; c++) {
            const size_t ch = (size_t)pair_of[o] + c;
            for (int a = 0; a < 3; a++) {
                nb[6 * (size_t)i + a] = std::min(nb[6 * (size_t)i + a], nb[6 * ch + a]);
                nb[6 * (size_t)i + 3 + a] = std::max(nb[6 * (size_t)i + 3 + a], nb[6 * ch + 3 + a]);
            }
        }
    }

    std::vector<float4> nodes((size_t)n_live * 2);
    parallel_ranges(n_live, nthr, [&](int n0, int n1) {
        for (int i = n0; i < n1; i++) {
            const int o = order[i];
            float x, y;
            if (T.axis[o] == -1) { x = as_float((int32_t)first_rec[i]); y = as_float_u(((uint32_t)leaf_recs_n[i] << 2) | WRT_LEAF_TAG); }
            else { x = T.split[o]; y = as_float_u(((uint32_t)pair_of[o] << 2) | (uint32_t)T.axis[o]); }
            const float* b = &nb[6 * (size_t)i];
            nodes[2 * (size_t)i] = make_float4(x, y, b[0], b[1]);
            nodes[2 * (size_t)i + 1] = make_float4(b[2], b[3], b[4], b[5]);
        }
    });

    std::vector<float4> prims((size_t)d->n_prims * 3);
    parallel_ranges(d->n_prims, nthr, [&](int i0, int i1) {
    for (int i = i0; i < i1; i++) {
        const float* pd = d->prim_data + 9 * (size_t)i;
        if (d->prim_kind[i] == WRT_PRIM_TRIANGLE) {
            prims[3 * (size_t)i] = make_float4(pd[0], pd[1], pd[2], as_float(d->prim_matid[i]));
            prims[3 * (size_t)i + 1] = make_float4(pd[3], pd[4], pd[5], as_float(WRT_PRIM_TRIANGLE));
            prims[3 * (size_t)i + 2] = make_float4(pd[6], pd[7], pd[8], 0.f);
        } else {
            prims[3 * (size_t)i] = make_float4(pd[0], pd[1], pd[2], as_float(d->prim_matid[i]));
            prims[3 * (size_t)i + 1] = make_float4(pd[3], 0.f, 0.f, as_float(WRT_PRIM_SPHERE));
            prims[3 * (size_t)i + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    });

    std::vector<DevMaterial> mats(std::max(d->n_materials, 1));
    memset(mats.data(), 0, mats.size() * sizeof(DevMaterial));
    for (int i = 0; i < d->n_materials; i++) {
        const float* m = d->materials + 11 * (size_t)i;
        for (int a = 0; a < 3; a++) { mats[i].diffuse[a] = m[a]; mats[i].phong[a] = m[3 + a]; mats[i].specular[a] = m[7 + a]; }
        mats[i].phong_exp = m[6]; mats[i].index = m[10];
    }
    std::vector<DevLight> lights(std::max(d->n_lights, 1));
    memset(lights.data(), 0, lights.size() * sizeof(DevLight));
    for (int i = 0; i < d->n_lights; i++) make_light(d->lights + 12 * (size_t)i, lights[i]);


    L.nodes.swap(nodes); L.recs.swap(recs); L.prims.swap(prims); L.materials.swap(mats); L.lights.swap(lights);
    L.n_nodes = n_live; L.n_recs = n_recs;
    DevSceneView& v = L.view;
    memset(&v, 0, sizeof v);
    v.small_tree = n_live < 512 ? 1 : 0;
    v.n_nodes = n_live; v.n_prims = d->n_prims; v.n_materials = d->n_materials; v.n_lights = d->n_lights;
    for (int a = 0; a < 3; a++) { v.root_lo[a] = T.root_box[a]; v.root_hi[a] = T.root_box[3 + a]; }
    {   // sceneSphere, scene.cpp:481-487
        float dx = v.root_hi[0] - v.root_lo[0], dy = v.root_hi[1] - v.root_lo[1], dz = v.root_hi[2] - v.root_lo[2];
        float diameter2 = dx * dx + dy * dy + dz * dz;
        for (int a = 0; a < 3; a++) v.sphere_center[a] = (v.root_lo[a] + v.root_hi[a]) * 0.5f;
        v.sphere_radius = std::sqrt(diameter2) * 0.5f;
        v.inv_sphere_radius_sqr = 1.f / diameter2;
    }
    return true;
}

void layout_point_view_at_host(SceneLayout& L)
{
    L.view.nodes = L.nodes.data(); L.view.leaf_recs = L.recs.data(); L.view.prims = L.prims.data();
    L.view.materials = L.materials.data(); L.view.lights = L.lights.data();
}

}  // namespace wrt
