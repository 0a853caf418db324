// SAH KD-tree construction that reproduces the reference's tree node for node.
//
// Restates KDtreeAccel::init / findSplitPlane / buildTree (R/src/scene/KDtreeAccel.cpp:12-307) and
// the per-primitive boxes of Triangle::setBox (R/src/geometry/triangle.h:15-23), Sphere::setBox
// (sphere.h:15-20) and AABB::extend (AABB.h:13-21).  Closest-hit selection in the reference is
// order dependent (KDtreeAccel.cpp:363-373), so bit-exact primitive ids need the SAME topology, the
// same leaf object order and the same split planes — including the effects of the epsilon-tolerant
// (non-transitive) event comparator under libc qsort (KDtreeAccel.cpp:3-10,41) and of never
// re-sorting the per-child event lists (:192-276).  The design differences are in the plumbing only:
// nodes are emitted straight into flat arrays (DFS pre-order), objects are indices not pointers, and
// a node's event lists are released as soon as its children own theirs (the reference keeps them all:
// ~2 GB at 1 M triangles).  The build is also parallel where that cannot change the result: the three initial
// event sorts run concurrently, the per-axis event distribution of large nodes runs on three threads, and the
// sub-trees below the top few levels are built as independent tasks into private arrays that are spliced back in
// DFS pre-order — the output arrays are identical, index for index, to the serial build's (tests/test_host.py).
//
// All float expressions keep the reference's operand order; this file must be compiled without FMA
// contraction (-ffp-contract=off, no -march=native).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <future>
#include <thread>
#include "host_scene.h"

namespace wrt {

namespace {

// Event::EventType values (KDtreeAccel.h:9-19): End sorts before Start at equal positions.
enum { kEnd = 0, kPlanar = 1, kStart = 2 };

struct Event {  // same 12-byte layout as the reference's Event so libc qsort sees the same problem
    Real pos;
    int type;
    int index;
};

int compare_events(const void* a, const void* b)  // cmp_sort_event, KDtreeAccel.cpp:3-10
{
    const Event* e1 = (const Event*)a;
    const Event* e2 = (const Event*)b;
    int c = cmp_eps(e1->pos - e2->pos);
    if (c != 0) return c;
    return e1->type - e2->type;
}

struct BuildNode {
    std::vector<int32_t> objs;     // indices into Scene::objs, in objlist order
    std::vector<Event> ev[3];
    Real lo[3], hi[3];             // node box (only meaningful when objs is non-empty)
};

inline Real surface_area(Real x, Real y, Real z)  // SA(), KDtreeAccel.cpp:59-62
{
    return 2 * (x * y + x * z + y * z);
}

// SAH(), KDtreeAccel.cpp:64-80
inline Real sah_cost(const BuildNode& nd, int axis, Real plane, int nl, int nr)
{
    Real v[3] = { nd.hi[0] - nd.lo[0], nd.hi[1] - nd.lo[1], nd.hi[2] - nd.lo[2] };
    Real vl[3] = { v[0], v[1], v[2] };
    Real vr[3] = { v[0], v[1], v[2] };
    vl[axis] = plane - nd.lo[axis];
    vr[axis] = nd.hi[axis] - plane;
    Real lambda = 1.0f;
    if (nl == 0 || nr == 0) lambda = 0.8f;
    return (lambda / surface_area(v[0], v[1], v[2])) *
           (surface_area(vl[0], vl[1], vl[2]) * nl + surface_area(vr[0], vr[1], vr[2]) * nr);
}

// findSplitPlane(), KDtreeAccel.cpp:82-116.  Returns axis (-1 if no candidate beat INF).
int find_split(const BuildNode& nd, Real* split_out)
{
    Real cost = kInf;
    int best_axis = -1;
    for (int axis = 0; axis < 3; axis++) {
        const std::vector<Event>& e = nd.ev[axis];
        const int n = (int)e.size();
        int nl = 0, nr = (int)nd.objs.size();
        int i = 0;
        while (i < n) {
            int p_end = 0, p_start = 0;
            Real now = e[i].pos;
            if (now != now) return -1;  // NaN coordinate: the reference would spin forever here
            while (i < n && e[i].pos == now) {
                if (e[i].type == kEnd) p_end++;
                if (e[i].type == kStart) p_start++;
                i++;
            }
            nr -= p_end;
            Real c = sah_cost(nd, axis, now, nl, nr);
            if (cmp_eps(c - cost) < 0) {
                cost = c;
                *split_out = now;
                best_axis = axis;
            }
            nl += p_start;
        }
    }
    return best_axis;
}

int emit_leaf(FlatTree& out, const BuildNode& nd, int dep)
{
    int me = (int)out.axis.size();
    out.axis.push_back(-1); out.split.push_back(0.f);
    out.left.push_back(-1); out.right.push_back(-1);
    out.first_ref.push_back((int32_t)out.refs.size());
    out.n_ref.push_back((int32_t)nd.objs.size());
    out.refs.insert(out.refs.end(), nd.objs.begin(), nd.objs.end());
    if (dep > out.depth) out.depth = dep;
    return me;
}

constexpr size_t kParallelAxisMin = 200000;   // events per axis above which the three axes are distributed concurrently

struct Builder {
    const std::vector<float>& boxes;  // 6 per prim
    int dep_max;
    int par_threads;                  // 1 = everything on the calling thread
    size_t chunk_min_events;          // events per axis above which an axis is distributed in several ranges

    // One node of buildTree(), KDtreeAccel.cpp:118-276: termination test, split search, classification and the
    // children's object / event lists.  Returns false when *nd stays a leaf; otherwise consumes nd's lists and
    // returns the children (nd itself is left empty, the caller deletes it).
    bool split_node(BuildNode* nd, int dep, int& axis, Real& split, BuildNode*& l, BuildNode*& r) const
    {
        if (dep > dep_max || nd->objs.size() <= 1) return false;
        split = 0.f;
        const bool prof = dep <= 2 && nd->objs.size() > 1000000 && getenv("WRT_KD_PROFILE");
        const auto tp0 = std::chrono::steady_clock::now();
        axis = find_split(*nd, &split);
        if (prof) fprintf(stderr, "[kd]   find_split %.2f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
        // No plane with cost < INF-EPS.  The reference indexes box.l[-1] here (undefined behaviour); the only
        // defined reading of the node it leaves is a leaf (axis == -1).
        if (axis < 0) return false;
        const int n = (int)nd->objs.size();
        enum { LeftOnly = 0, RightOnly = 1, Both = 2 };
        std::vector<unsigned char> div(n);
        std::vector<int32_t> to_l(n), to_r(n);
        l = new BuildNode();
        r = new BuildNode();
        int nl = 0, nr = 0, nb = 0;
        for (int i = 0; i < n; i++) {  // :142-163
            const float* b = &boxes[6 * (size_t)nd->objs[i]];
            Real st = b[axis], ed = b[3 + axis];
            if (cmp_eps(ed - split) <= 0) { div[i] = LeftOnly; nl++; }
            else if (cmp_eps(split - st) <= 0) { div[i] = RightOnly; nr++; }
            else { div[i] = Both; nb++; }
        }
        l->objs.reserve(nl + nb); r->objs.reserve(nb + nr);
        for (int i = 0; i < n; i++) {  // :180-201
            if (div[i] == LeftOnly) { to_l[i] = (int)l->objs.size(); l->objs.push_back(nd->objs[i]); }
            else if (div[i] == RightOnly) { to_r[i] = (int)r->objs.size(); r->objs.push_back(nd->objs[i]); }
            else {
                to_l[i] = (int)l->objs.size(); l->objs.push_back(nd->objs[i]);
                to_r[i] = (int)r->objs.size(); r->objs.push_back(nd->objs[i]);
            }
        }
        if (prof) fprintf(stderr, "[kd]   + classify %.2f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
        const int split_axis = axis;
        const Real split_pos = split;
        // :203-276; children inherit the parent's order, never re-sorted.  One event range [j0, j1) of axis `a` into the
        // given child lists (the lists of an axis are filled in event order, so ranges can be processed independently
        // and concatenated).
        auto distribute_range = [&](int a, size_t j0, size_t j1, std::vector<Event>& le, std::vector<Event>& re) {
            const std::vector<Event>& pe = nd->ev[a];
            for (size_t j = j0; j < j1; j++) {
                const Event& s = pe[j];
                Event e;
                e.type = s.type;
                const int d = div[s.index];
                if (d == LeftOnly) { e.pos = s.pos; e.index = to_l[s.index]; le.push_back(e); }
                else if (d == RightOnly) { e.pos = s.pos; e.index = to_r[s.index]; re.push_back(e); }
                else if (a != split_axis) {
                    e.pos = s.pos;
                    e.index = to_l[s.index]; le.push_back(e);
                    e.index = to_r[s.index]; re.push_back(e);
                } else if (s.type == kEnd) {     // straddler's end: clipped to the plane on the left
                    e.pos = split_pos; e.index = to_l[s.index]; le.push_back(e);
                    e.pos = s.pos; e.index = to_r[s.index]; re.push_back(e);
                } else if (s.type == kStart) {   // straddler's start: clipped on the right
                    e.pos = s.pos; e.index = to_l[s.index]; le.push_back(e);
                    e.pos = split_pos; e.index = to_r[s.index]; re.push_back(e);
                }
            }
        };
        const size_t n_ev = nd->ev[0].size();
        const int chunks = (n_ev >= chunk_min_events && par_threads >= 6) ? std::min(par_threads / 3, 8) : 1;
        auto distribute = [&](int a) {
            std::vector<Event>& pe = nd->ev[a];
            l->ev[a].reserve(l->objs.size() * 2);
            r->ev[a].reserve(r->objs.size() * 2);
            if (chunks <= 1) distribute_range(a, 0, pe.size(), l->ev[a], r->ev[a]);
            else {   // large node: the axis in `chunks` ranges on their own threads, concatenated in order
                std::vector<std::vector<Event>> lp(chunks), rp(chunks);
                std::vector<std::thread> th;
                for (int c = 0; c < chunks; c++) {
                    const size_t j0 = pe.size() * c / chunks, j1 = pe.size() * (c + 1) / chunks;
                    lp[c].reserve((j1 - j0)); rp[c].reserve((j1 - j0));
                    th.emplace_back([&, a, c, j0, j1] { distribute_range(a, j0, j1, lp[c], rp[c]); });
                }
                for (auto& t : th) t.join();
                for (int c = 0; c < chunks; c++) {
                    l->ev[a].insert(l->ev[a].end(), lp[c].begin(), lp[c].end());
                    r->ev[a].insert(r->ev[a].end(), rp[c].begin(), rp[c].end());
                    std::vector<Event>().swap(lp[c]); std::vector<Event>().swap(rp[c]);
                }
            }
            std::vector<Event>().swap(pe);  // parent's list is no longer needed
        };
        if (nd->ev[0].size() >= kParallelAxisMin && par_threads > 1) {   // the axes are independent: same lists, three threads
            std::thread t1(distribute, 1), t2(distribute, 2);
            distribute(0);
            t1.join(); t2.join();
        } else {
            for (int a = 0; a < 3; a++) distribute(a);
        }
        BuildNode* kids[2] = { l, r };
        for (int k = 0; k < 2; k++) {  // :277-294 — child box = first/last event per axis
            BuildNode* c = kids[k];
            if (!c->objs.empty())
                for (int a = 0; a < 3; a++) { c->lo[a] = c->ev[a].front().pos; c->hi[a] = c->ev[a].back().pos; }
            else
                for (int a = 0; a < 3; a++) { c->lo[a] = 0.f; c->hi[a] = 0.f; }
        }
        std::vector<int32_t>().swap(nd->objs);
        return true;
    }

    // buildTree(), KDtreeAccel.cpp:118-307, serial: nodes are appended to `out` in DFS pre-order.  Consumes *nd.
    int build(BuildNode* nd, int dep, FlatTree& out) const
    {
        int axis; Real split; BuildNode* l; BuildNode* r;
        const int n = (int)nd->objs.size();
        if (!split_node(nd, dep, axis, split, l, r)) {
            int me = emit_leaf(out, *nd, dep);
            delete nd;
            return me;
        }
        delete nd;
        int me = (int)out.axis.size();
        out.axis.push_back(axis); out.split.push_back(split);
        out.left.push_back(-1); out.right.push_back(-1);
        out.first_ref.push_back(-1); out.n_ref.push_back(n);
        int li = build(l, dep + 1, out);
        int ri = build(r, dep + 1, out);
        out.left[me] = li; out.right[me] = ri;
        return me;
    }

    // Appends sub-tree `t` (its own indices start at 0) to `out`; returns the index its root gets.
    static int splice(FlatTree& out, const FlatTree& t)
    {
        const int32_t node_off = (int32_t)out.axis.size();
        const int32_t ref_off = (int32_t)out.refs.size();
        out.axis.insert(out.axis.end(), t.axis.begin(), t.axis.end());
        out.split.insert(out.split.end(), t.split.begin(), t.split.end());
        out.n_ref.insert(out.n_ref.end(), t.n_ref.begin(), t.n_ref.end());
        for (size_t i = 0; i < t.axis.size(); i++) {
            out.left.push_back(t.left[i] < 0 ? -1 : t.left[i] + node_off);
            out.right.push_back(t.right[i] < 0 ? -1 : t.right[i] + node_off);
            out.first_ref.push_back(t.first_ref[i] < 0 ? -1 : t.first_ref[i] + ref_off);
        }
        out.refs.insert(out.refs.end(), t.refs.begin(), t.refs.end());
        if (t.depth > out.depth) out.depth = t.depth;
        return node_off;
    }

    // Same tree, the two sub-trees of the top `levels` levels built concurrently into private arrays and spliced
    // back in pre-order (node, left sub-tree, right sub-tree): identical output to build().
    void build_parallel(BuildNode* nd, int dep, FlatTree& out, int levels) const
    {
        if (levels <= 0 || nd->objs.size() < 4096) { build(nd, dep, out); return; }
        int axis; Real split; BuildNode* l; BuildNode* r;
        const int n = (int)nd->objs.size();
        const auto t0 = std::chrono::steady_clock::now();
        const bool ok = split_node(nd, dep, axis, split, l, r);
        if (dep <= 3 && getenv("WRT_KD_PROFILE"))
            fprintf(stderr, "[kd] split at depth %d, %d objects: %.2f s\n", dep, n,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        if (!ok) {
            emit_leaf(out, *nd, dep);
            delete nd;
            return;
        }
        delete nd;
        FlatTree lt, rt;
        lt.depth = rt.depth = 0;
        std::future<void> fl = std::async(std::launch::async, [&] { build_parallel(l, dep + 1, lt, levels - 1); });
        build_parallel(r, dep + 1, rt, levels - 1);
        fl.get();
        const int me = (int)out.axis.size();
        out.axis.push_back(axis); out.split.push_back(split);
        out.left.push_back(-1); out.right.push_back(-1);
        out.first_ref.push_back(-1); out.n_ref.push_back(n);
        const int li = splice(out, lt);
                const int ri = splice(out, rt);
        out.left[me] = li; out.right[me] = ri;
    }
};

inline Real min2(Real a, Real b) { return (b < a) ? b : a; }  // std::min
inline Real max2(Real a, Real b) { return (a < b) ? b : a; }  // std::max

}  // namespace

void prim_box(int kind, const float* d, float b[6])
{
    if (kind == WRT_PRIM_TRIANGLE) {  // Triangle::setBox, triangle.h:15-23
        for (int a = 0; a < 3; a++) {
            b[a] = min2(d[a], min2(d[3 + a], d[6 + a]));
            b[3 + a] = max2(d[a], max2(d[3 + a], d[6 + a]));
        }
    } else {  // Sphere::setBox, sphere.h:15-20
        for (int a = 0; a < 3; a++) { b[a] = d[a] - d[3]; b[3 + a] = d[a] + d[3]; }
    }
    for (int a = 0; a < 3; a++)  // AABB::extend, AABB.h:13-21 ("avoid 2-D box")
        if (cmp_eps(b[a] - b[3 + a]) == 0) b[3 + a] += 10 * kEps;
}

bool build_kdtree(HostScene& hs, std::string& err)
{
    const int n = hs.n_prims();
    hs.tree.clear();
    hs.tree_built = false;
    if (n <= 0) { err = "build_kdtree: scene has no primitives"; return false; }
    std::vector<float> boxes((size_t)n * 6);
    for (int i = 0; i < n; i++) prim_box(hs.prim_kind[i], &hs.prim_data[9 * (size_t)i], &boxes[6 * (size_t)i]);

    // WRT_KD_THREADS=1 forces the serial build (the tests compare the two)
    int threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("WRT_KD_THREADS")) threads = atoi(e);
    if (threads < 1) threads = 1;

    // KDtreeAccel::init, KDtreeAccel.cpp:12-57
    hs.tree.dep_max = (int)(1.2 * std::log((double)n) + 2.0);
    BuildNode* root = new BuildNode();
    root->objs.resize(n);
    for (int i = 0; i < n; i++) root->objs[i] = i;
    for (int a = 0; a < 3; a++) {
        std::vector<Event>& e = root->ev[a];
        e.resize((size_t)2 * n);
        for (int j = 0; j < n; j++) {
            e[2 * (size_t)j].type = kStart; e[2 * (size_t)j].pos = boxes[6 * (size_t)j + a]; e[2 * (size_t)j].index = j;
            e[2 * (size_t)j + 1].type = kEnd; e[2 * (size_t)j + 1].pos = boxes[6 * (size_t)j + 3 + a]; e[2 * (size_t)j + 1].index = j;
        }
    }
    const bool profile = getenv("WRT_KD_PROFILE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    {   // libc qsort on purpose: the comparator is not a strict weak order, so the result depends on
        // the sorting algorithm; the reference's tree is whatever libc qsort makes of it.  The three arrays are
        // independent, so they are sorted concurrently.
        auto sort_axis = [&](int a) { qsort(root->ev[a].data(), root->ev[a].size(), sizeof(Event), compare_events); };
        if (threads > 1) { std::thread t1(sort_axis, 1), t2(sort_axis, 2); sort_axis(0); t1.join(); t2.join(); }
        else for (int a = 0; a < 3; a++) sort_axis(a);
    }
    if (profile) fprintf(stderr, "[kd] events + sort: %.2f s\n", since(t_start));
    const auto t_build = std::chrono::steady_clock::now();
    for (int a = 0; a < 3; a++) {
        root->lo[a] = root->ev[a].front().pos;
        root->hi[a] = root->ev[a].back().pos;
    }
    for (int a = 0; a < 3; a++) { hs.tree.root_box[a] = root->lo[a]; hs.tree.root_box[3 + a] = root->hi[a]; }

    size_t chunk_min = 4 * kParallelAxisMin;
    if (const char* e = getenv("WRT_KD_CHUNK_MIN")) chunk_min = (size_t)atoll(e);     // tests lower it to exercise the path
    Builder b = { boxes, hs.tree.dep_max, threads, chunk_min };
    if (threads > 1) {
        int levels = 1;
        while ((1 << levels) < 2 * threads && levels < 7) levels++;   // ~2 tasks per thread
        b.build_parallel(root, 1, hs.tree, levels);
    } else {
        b.build(root, 1, hs.tree);
    }
    hs.tree_built = true;
    if (profile) fprintf(stderr, "[kd] build: %.2f s (%d threads), %zu nodes\n", since(t_build), threads, hs.tree.axis.size());

    // sceneSphere, scene.cpp:481-487
    const float* rb = hs.tree.root_box;
    Real dx = rb[3] - rb[0], dy = rb[4] - rb[1], dz = rb[5] - rb[2];
    Real diameter2 = dx * dx + dy * dy + dz * dz;
    hs.scene_sphere[0] = (rb[0] + rb[3]) * 0.5f;
    hs.scene_sphere[1] = (rb[1] + rb[4]) * 0.5f;
    hs.scene_sphere[2] = (rb[2] + rb[5]) * 0.5f;
    hs.scene_sphere[3] = std::sqrt(diameter2) * 0.5f;
    hs.scene_sphere[4] = 1.f / diameter2;
    return true;
}

}  // namespace wrt
