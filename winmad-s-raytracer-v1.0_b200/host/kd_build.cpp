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
// ~2 GB at 1 M triangles).  The build is also parallel where that cannot change the result (round 2: every phase of the
// large top nodes):
//   * the initial event sorts: glibc >= 2.39's qsort IS a top-down merge sort (stdlib/qsort.c msort_with_tmp: halves
//     n/2 | n - n/2, the merge takes the left element when cmp <= 0), so the same recursion with the halves sorted on
//     separate threads performs exactly the same comparisons with the same outcome whatever the comparator — no
//     transitivity needed.  A probe at first use checks the replica against the process's libc qsort on tie-heavy data
//     and falls back to libc qsort if they differ (another libc, another algorithm: the reference's tree is whatever
//     ITS libc makes of the comparator);
//   * findSplitPlane: a candidate can only be accepted (`cost - EPS` rule) if it is a strict prefix minimum of its
//     axis' cost sequence, so each axis / each range of an axis reports its short staircase of prefix minima and the
//     reference's running-minimum rule is replayed over the concatenated staircases in order (find_split_parallel);
//   * classification, child object lists and the per-axis event distribution in ranges with counted offsets;
//   * the sub-trees below the top few levels as independent tasks into private arrays spliced back in DFS pre-order.
// The output arrays are identical, index for index, to the serial build's and to the reference's (tests/test_host.py).
//
// All float expressions keep the reference's operand order; this file must be compiled without FMA
// contraction (-ffp-contract=off, no -march=native).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <atomic>
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

// SAH() with everything that does not depend on the candidate evaluated once per (node, axis): the two possible values of
// lambda / SA(v) and the products of the two extents the plane does not move.  Every float operation of sah_cost() is still
// performed on the same operands in the same order (x*y + x*z + y*z, left to right), so the costs are bit-identical.
struct SahAxis {
    int axis;
    Real lo, hi, q1, q08;      // q = lambda / SA(v)
    Real u, w, uw;             // the other two extents (in x, y, z order) and their product
    explicit SahAxis(const BuildNode& nd, int a) : axis(a), lo(nd.lo[a]), hi(nd.hi[a])
    {
        const Real v[3] = { nd.hi[0] - nd.lo[0], nd.hi[1] - nd.lo[1], nd.hi[2] - nd.lo[2] };
        const Real sa = surface_area(v[0], v[1], v[2]);
        q1 = 1.0f / sa; q08 = 0.8f / sa;
        u = v[a == 0 ? 1 : 0]; w = v[a == 2 ? 1 : 2]; uw = u * w;
    }
    inline Real sa_with(Real t) const      // SA of the node box with extent t along `axis`
    {
        // axis 0: x=t,y=u,z=w: t*u + t*w + u*w;  axis 1: x=u,y=t,z=w: u*t + u*w + t*w;  axis 2: x=u,y=w,z=t: u*w + u*t + w*t
        if (axis == 0) return 2 * (t * u + t * w + uw);
        if (axis == 1) return 2 * (u * t + uw + t * w);
        return 2 * (uw + u * t + w * t);
    }
    inline Real cost(Real plane, int nl, int nr) const
    {
        const Real q = (nl == 0 || nr == 0) ? q08 : q1;
        return q * (sa_with(plane - lo) * nl + sa_with(hi - plane) * nr);
    }
};

// findSplitPlane(), KDtreeAccel.cpp:82-116.  Returns axis (-1 if no candidate beat INF).
int find_split(const BuildNode& nd, Real* split_out)
{
    Real cost = kInf;
    int best_axis = -1;
    for (int axis = 0; axis < 3; axis++) {
        const std::vector<Event>& e = nd.ev[axis];
        const int n = (int)e.size();
        const SahAxis sah(nd, axis);
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
            Real c = sah.cost(now, nl, nr);
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


template <class F>
void parallel_tasks(int n_tasks, F fn)      // fn(task); task 0 runs on the calling thread
{
    std::vector<std::thread> th;
    for (int t = 1; t < n_tasks; t++) th.emplace_back([&fn, t] { fn(t); });
    fn(0);
    for (auto& x : th) x.join();
}

// ---- glibc's merge sort, restated (see the header comment) -------------------------------------------
void msort_serial(Event* b, size_t n, Event* tmp)
{
    if (n <= 1) return;
    size_t n1 = n / 2, n2 = n - n1;
    Event* b1 = b; Event* b2 = b + n1;
    msort_serial(b1, n1, tmp);
    msort_serial(b2, n2, tmp);
    Event* t = tmp;
    while (n1 > 0 && n2 > 0) {
        if (compare_events(b1, b2) <= 0) { *t++ = *b1++; --n1; }
        else { *t++ = *b2++; --n2; }
    }
    if (n1 > 0) memcpy(t, b1, n1 * sizeof(Event));
    memcpy(b, tmp, (n - n2) * sizeof(Event));
}

// Same recursion; the two halves of the top `levels` levels on their own threads, each with its own slice of tmp.
void msort_parallel(Event* b, size_t n, Event* tmp, int levels)
{
    if (levels <= 0 || n < 32768) { msort_serial(b, n, tmp); return; }
    size_t n1 = n / 2, n2 = n - n1;
    Event* b1 = b; Event* b2 = b + n1;
    std::thread th([=] { msort_parallel(b1, n1, tmp, levels - 1); });
    msort_parallel(b2, n2, tmp + n1, levels - 1);
    th.join();
    Event* t = tmp;
    while (n1 > 0 && n2 > 0) {
        if (compare_events(b1, b2) <= 0) { *t++ = *b1++; --n1; }
        else { *t++ = *b2++; --n2; }
    }
    if (n1 > 0) memcpy(t, b1, n1 * sizeof(Event));
    memcpy(b, tmp, (n - n2) * sizeof(Event));
}

// Does this process's libc qsort order tie-heavy events exactly like the restated merge sort?  (Checked once.)
bool libc_qsort_is_msort()
{
    static const bool same = [] {
        for (size_t n : { (size_t)7, (size_t)1000, (size_t)40001 }) {
            std::vector<Event> a(n), b, tmp(n);
            uint32_t x = 12345u;
            for (size_t i = 0; i < n; i++) {
                x = x * 1664525u + 1013904223u;
                a[i].pos = (float)(x >> 8) * (1.0f / 16777216.0f) * (n < 2000 ? 0.02f : 1.5f);    // many EPS-ties
                a[i].type = (i & 1) ? kEnd : kStart; a[i].index = (int)(i / 2);
            }
            b = a;
            qsort(a.data(), n, sizeof(Event), compare_events);
            msort_parallel(b.data(), n, tmp.data(), 2);
            if (memcmp(a.data(), b.data(), n * sizeof(Event)) != 0) return false;
        }
        return true;
    }();
    return same;
}

void sort_events(std::vector<Event>& e, int threads)
{
    const char* force = getenv("WRT_KD_SORT");      // "libc": always libc qsort (the tests compare the two)
    size_t min_n = 65536;                           // below this one thread is as fast
    if (const char* m = getenv("WRT_KD_SORT_MIN")) min_n = (size_t)atoll(m);      // (tests lower it)
    if (threads <= 1 || e.size() < min_n || (force && force[0] == 'l') || !libc_qsort_is_msort()) {
        qsort(e.data(), e.size(), sizeof(Event), compare_events);
        return;
    }
    int levels = 0;
    while ((1 << levels) < threads && levels < 6) levels++;
    std::vector<Event> tmp(e.size());
    msort_parallel(e.data(), e.size(), tmp.data(), levels);
}

// findSplitPlane() for large nodes, exactly (see the header comment).  An axis is cut into ranges at position-group
// boundaries; a range reports (a) its Start / End counts, then, knowing nl / nr at its first group, (b) the strict prefix
// minima of its candidates' costs.  A candidate that is not a strict prefix minimum of its own axis can never satisfy
// `c - cost < -EPS`: every earlier candidate c' of the axis was seen with cost' >= cost and left cost' - EPS <= c' or became
// the cost, so c < cost - EPS implies c < c'.  Replaying the rule over the staircases, axis by axis in order, therefore
// makes the same decisions as the full scan.
struct SplitCand { Real c, pos; };

int find_split_parallel(const BuildNode& nd, Real* split_out, int threads)
{
    const int per_axis = std::max(1, threads / 3);
    struct Range { size_t j0, j1; int starts, ends; bool nan; std::vector<SplitCand> stairs; };
    std::vector<Range> R((size_t)3 * per_axis);
    for (int a = 0; a < 3; a++) {
        const std::vector<Event>& e = nd.ev[a];
        const size_t n = e.size();
        size_t prev = 0;
        for (int k = 0; k < per_axis; k++) {
            size_t j1 = (k + 1 == per_axis) ? n : n * (size_t)(k + 1) / per_axis;
            if (j1 < prev) j1 = prev;
            while (j1 > prev && j1 < n && e[j1].pos == e[j1 - 1].pos) j1++;     // never cut a group of equal positions
            Range& r = R[(size_t)a * per_axis + k];
            r.j0 = prev; r.j1 = j1; r.starts = r.ends = 0; r.nan = false;
            prev = j1;
        }
    }
    parallel_tasks(3 * per_axis, [&](int t) {
        Range& r = R[t];
        const std::vector<Event>& e = nd.ev[t / per_axis];
        int s = 0, en = 0;
        for (size_t j = r.j0; j < r.j1; j++) { s += (e[j].type == kStart); en += (e[j].type == kEnd); }
        r.starts = s; r.ends = en;
    });
    parallel_tasks(3 * per_axis, [&](int t) {
        Range& r = R[t];
        const int axis = t / per_axis;
        const std::vector<Event>& e = nd.ev[axis];
        const SahAxis sah(nd, axis);
        int nl = 0, nr = (int)nd.objs.size();
        for (int k = axis * per_axis; k < t; k++) { nl += R[k].starts; nr -= R[k].ends; }
        Real run = 0.f; bool have = false;
        size_t i = r.j0;
        while (i < r.j1) {
            int p_end = 0, p_start = 0;
            const Real now = e[i].pos;
            if (now != now) { r.nan = true; return; }
            while (i < r.j1 && e[i].pos == now) {
                if (e[i].type == kEnd) p_end++;
                if (e[i].type == kStart) p_start++;
                i++;
            }
            nr -= p_end;
            const Real c = sah.cost(now, nl, nr);
            if (!have ? (c == c) : (c < run)) { run = c; have = true; r.stairs.push_back({ c, now }); }
            nl += p_start;
        }
    });
    Real cost = kInf;
    int best_axis = -1;
    for (size_t t = 0; t < R.size(); t++) {
        if (R[t].nan) return -1;     // the serial scan returns at the first NaN it meets; a NaN position poisons the node either way
        for (const SplitCand& sc : R[t].stairs)
            if (cmp_eps(sc.c - cost) < 0) { cost = sc.c; *split_out = sc.pos; best_axis = (int)(t / per_axis); }
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

// WRT_KD_PROFILE: thread-summed time per phase of split_node (nanoseconds)
std::atomic<long long> g_ns_find{0}, g_ns_classify{0}, g_ns_distribute{0}, g_ns_leaf{0};
const bool g_kd_profile = getenv("WRT_KD_PROFILE") != nullptr;
inline long long now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

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
        const bool big = par_threads >= 3 && nd->ev[0].size() >= chunk_min_events;      // a top node: every phase in ranges on threads
        const long long pn0 = g_kd_profile ? now_ns() : 0;
        axis = big ? find_split_parallel(*nd, &split, par_threads) : find_split(*nd, &split);
        const long long pn1 = g_kd_profile ? now_ns() : 0;
        if (g_kd_profile) g_ns_find += pn1 - pn0;
        if (prof) fprintf(stderr, "[kd]   find_split %.2f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
        // No plane with cost < INF-EPS.  The reference indexes box.l[-1] here (undefined behaviour); the only
        // defined reading of the node it leaves is a leaf (axis == -1).
        if (axis < 0) return false;
        const int n = (int)nd->objs.size();
        // per object: its index in the left / right child's list, -1 where it does not go (one 8-byte record = one cache line
        // touched per event in the distribution below, instead of three arrays)
        struct LR { int32_t l, r; };
        std::vector<LR> lr(n);
        l = new BuildNode();
        r = new BuildNode();
        // classification (:142-163) and the children's object lists (:180-201): object ranges with counted offsets
        const int oc = big ? par_threads : 1;
        std::vector<int> cnt_l(oc + 1, 0), cnt_r(oc + 1, 0);     // objects a range sends left / right (straddlers count in both)
        parallel_tasks(oc, [&](int t) {
            const int i0 = (int)((long long)n * t / oc), i1 = (int)((long long)n * (t + 1) / oc);
            int cl = 0, cr = 0;
            for (int i = i0; i < i1; i++) {
                const float* b = &boxes[6 * (size_t)nd->objs[i]];
                Real st = b[axis], ed = b[3 + axis];
                if (cmp_eps(ed - split) <= 0) { lr[i].l = 0; lr[i].r = -1; cl++; }            // LeftOnly
                else if (cmp_eps(split - st) <= 0) { lr[i].l = -1; lr[i].r = 0; cr++; }    // RightOnly
                else { lr[i].l = 0; lr[i].r = 0; cl++; cr++; }                               // Both
            }
            cnt_l[t + 1] = cl; cnt_r[t + 1] = cr;
        });
        for (int t = 0; t < oc; t++) { cnt_l[t + 1] += cnt_l[t]; cnt_r[t + 1] += cnt_r[t]; }
        l->objs.resize(cnt_l[oc]); r->objs.resize(cnt_r[oc]);
        parallel_tasks(oc, [&](int t) {
            const int i0 = (int)((long long)n * t / oc), i1 = (int)((long long)n * (t + 1) / oc);
            int pl = cnt_l[t], pr = cnt_r[t];
            for (int i = i0; i < i1; i++) {
                const int32_t o = nd->objs[i];
                if (lr[i].l >= 0) { lr[i].l = pl; l->objs[pl++] = o; }
                if (lr[i].r >= 0) { lr[i].r = pr; r->objs[pr++] = o; }
            }
        });
        if (prof) fprintf(stderr, "[kd]   + classify %.2f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
        const long long pn2 = g_kd_profile ? now_ns() : 0;
        if (g_kd_profile) g_ns_classify += pn2 - pn1;
        const int split_axis = axis;
        const Real split_pos = split;
        // :203-276; children inherit the parent's order, never re-sorted.  One event range [j0, j1) of axis `a` written to
        // le / re (the lists of an axis are filled in event order, so ranges are independent once their offsets are known).
        auto count_range = [&](int a, size_t j0, size_t j1, size_t& nle, size_t& nre) {
            const std::vector<Event>& pe = nd->ev[a];
            size_t cl = 0, cr = 0;
            for (size_t j = j0; j < j1; j++) {
                const LR d = lr[pe[j].index];
                if (d.r < 0) cl++;
                else if (d.l < 0) cr++;
                else if (a != split_axis || pe[j].type == kEnd || pe[j].type == kStart) { cl++; cr++; }
            }
            nle = cl; nre = cr;
        };
        auto distribute_range = [&](int a, size_t j0, size_t j1, Event* le, Event* re) -> std::pair<size_t, size_t> {
            const std::vector<Event>& pe = nd->ev[a];
            Event* const le0 = le; Event* const re0 = re;
            for (size_t j = j0; j < j1; j++) {
                const Event& s = pe[j];
                Event e;
                e.type = s.type;
                const LR d = lr[s.index];
                if (d.r < 0) { e.pos = s.pos; e.index = d.l; *le++ = e; }            // LeftOnly
                else if (d.l < 0) { e.pos = s.pos; e.index = d.r; *re++ = e; }       // RightOnly
                else if (a != split_axis) {
                    e.pos = s.pos;
                    e.index = d.l; *le++ = e;
                    e.index = d.r; *re++ = e;
                } else if (s.type == kEnd) {     // straddler's end: clipped to the plane on the left
                    e.pos = split_pos; e.index = d.l; *le++ = e;
                    e.pos = s.pos; e.index = d.r; *re++ = e;
                } else if (s.type == kStart) {   // straddler's start: clipped on the right
                    e.pos = s.pos; e.index = d.l; *le++ = e;
                    e.pos = split_pos; e.index = d.r; *re++ = e;
                }
            }
            return { (size_t)(le - le0), (size_t)(re - re0) };
        };
        const int chunks = big ? std::max(1, par_threads / 3) : 1;
        if (chunks > 1) {
            // large node: 3 axes x `chunks` ranges; pass 1 counts what each range sends to either child, pass 2 writes at the offsets
            const int nt = 3 * chunks;
            std::vector<size_t> off_l(nt + 3, 0), off_r(nt + 3, 0);
            parallel_tasks(nt, [&](int t) {
                const int a = t / chunks, c = t % chunks;
                const size_t m = nd->ev[a].size();
                count_range(a, m * c / chunks, m * (c + 1) / chunks, off_l[t], off_r[t]);
            });
            for (int a = 0; a < 3; a++) {
                size_t sl = 0, sr = 0;
                for (int c = 0; c < chunks; c++) { const int t = a * chunks + c; const size_t cl = off_l[t], cr = off_r[t]; off_l[t] = sl; off_r[t] = sr; sl += cl; sr += cr; }
                l->ev[a].resize(sl); r->ev[a].resize(sr);
            }
            parallel_tasks(nt, [&](int t) {
                const int a = t / chunks, c = t % chunks;
                const size_t m = nd->ev[a].size();
                distribute_range(a, m * c / chunks, m * (c + 1) / chunks, l->ev[a].data() + off_l[t], r->ev[a].data() + off_r[t]);
            });
            for (int a = 0; a < 3; a++) std::vector<Event>().swap(nd->ev[a]);
        } else {
            auto distribute = [&](int a) {
                std::vector<Event>& pe = nd->ev[a];
                // every object sends exactly two events per axis to each child that holds it (a straddler's clipped pair included)
                l->ev[a].resize(l->objs.size() * 2);
                r->ev[a].resize(r->objs.size() * 2);
                const auto w = distribute_range(a, 0, pe.size(), l->ev[a].data(), r->ev[a].data());
                l->ev[a].resize(w.first); r->ev[a].resize(w.second);
                std::vector<Event>().swap(pe);  // parent's list is no longer needed
            };
            if (nd->ev[0].size() >= kParallelAxisMin && par_threads > 1) {   // the axes are independent: same lists, three threads
                std::thread t1(distribute, 1), t2(distribute, 2);
                distribute(0);
                t1.join(); t2.join();
            } else {
                for (int a = 0; a < 3; a++) distribute(a);
            }
        }
        if (g_kd_profile) g_ns_distribute += now_ns() - pn2;
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

    // Same tree, the two sub-trees of every node with more than `task_objs` objects built concurrently into private arrays and
    // spliced back in pre-order (node, left sub-tree, right sub-tree): identical output to build().  Tasks are cut by SIZE, not by
    // depth: the reference's SAH peels thin slabs off the big scenes first (C5: 112 482 | 10 102 328 objects at the root), and a
    // depth rule would spend its levels on those.
    void build_parallel(BuildNode* nd, int dep, FlatTree& out, size_t task_objs) const
    {
        if (nd->objs.size() <= task_objs || nd->objs.size() < 4096) { build(nd, dep, out); return; }
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
        std::future<void> fl = std::async(std::launch::async, [&] { build_parallel(l, dep + 1, lt, task_objs); });
        build_parallel(r, dep + 1, rt, task_objs);
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
    // WRT_KD_THREADS=1 forces the serial build (the tests compare the two)
    int threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("WRT_KD_THREADS")) threads = atoi(e);
    if (threads < 1) threads = 1;
    {
        const int bt = n >= 100000 ? threads : 1;
        parallel_tasks(bt, [&](int t) {
            const int i0 = (int)((long long)n * t / bt), i1 = (int)((long long)n * (t + 1) / bt);
            for (int i = i0; i < i1; i++) prim_box(hs.prim_kind[i], &hs.prim_data[9 * (size_t)i], &boxes[6 * (size_t)i]);
        });
    }

    // KDtreeAccel::init, KDtreeAccel.cpp:12-57
    hs.tree.dep_max = (int)(1.2 * std::log((double)n) + 2.0);
    BuildNode* root = new BuildNode();
    root->objs.resize(n);
    for (int i = 0; i < n; i++) root->objs[i] = i;
    parallel_tasks(threads > 1 ? 3 : 1, [&](int t) {
        for (int a = (threads > 1 ? t : 0); a < (threads > 1 ? t + 1 : 3); a++) {
            std::vector<Event>& e = root->ev[a];
            e.resize((size_t)2 * n);
            for (int j = 0; j < n; j++) {
                e[2 * (size_t)j].type = kStart; e[2 * (size_t)j].pos = boxes[6 * (size_t)j + a]; e[2 * (size_t)j].index = j;
                e[2 * (size_t)j + 1].type = kEnd; e[2 * (size_t)j + 1].pos = boxes[6 * (size_t)j + 3 + a]; e[2 * (size_t)j + 1].index = j;
            }
        }
    });
    const bool profile = getenv("WRT_KD_PROFILE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    {   // The comparator is not a strict weak order, so the result depends on the sorting algorithm; the reference's tree is
        // whatever libc qsort makes of it: sort_events() is libc qsort or its exact multi-threaded restatement.  The three
        // arrays are independent, so they are sorted concurrently.
        const int per_axis = threads > 1 ? std::max(2, (threads + 1) / 2) : 1;      // the merges are memory-bound: mild over-subscription
        auto sort_axis = [&](int a) { sort_events(root->ev[a], per_axis); };
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
        size_t task_objs = (size_t)n / (size_t)(4 * threads) + 1;     // ~4 tasks per thread by size
        if (const char* e = getenv("WRT_KD_TASK_OBJS")) task_objs = (size_t)atoll(e);
        b.build_parallel(root, 1, hs.tree, task_objs);
    } else {
        b.build(root, 1, hs.tree);
    }
    hs.tree_built = true;
    if (profile) fprintf(stderr, "[kd] build: %.2f s (%d threads), %zu nodes; thread-summed: find_split %.2f s, classify %.2f s, distribute %.2f s\n", since(t_build), threads, hs.tree.axis.size(),
                         g_ns_find.load() * 1e-9, g_ns_classify.load() * 1e-9, g_ns_distribute.load() * 1e-9);

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
