// Pooled traversal (scheduler 3): a warp owns a pool of kPoolRays rays whose traversal state lives in
// shared memory, and three FIFO rings of slot ids — rays waiting for a node visit, rays waiting for
// primitive tests, free slots.  Each round the warp serves the longer ring: lane l takes the l-th
// waiting ray, loads its state (a few LDS.128), performs one unit of work (<= kNodeSteps = 6 node visits or
// <= kPrimSteps = 4 leaf records, skip records included) and re-queues the ray on the ring it now belongs to
// (ballot/popc positions).  Because there are twice as many rays as lanes, rounds start with most lanes busy
// whatever the length of the individual rays (they differ by 50x in this tree; 3 of 32 lanes were busy when every
// lane was married to one ray, profiles/r1_ncu_extend_*).  Free slots are refilled from the global work counter
// in batches of >= 32.  Tuning of every constant here: profiles/r1_experiments.md.
//
// The per-ray steps are exactly those of traverse.cuh / trace_persistent.cuh (trav_begin, trav_skip,
// trav_interior arithmetic, leaf_record, pop), executed in the same order for every ray, so results are
// bit-identical to kd_traverse().  Only the traversal stack moves: it lives in a per-warp global
// scratch area ([depth][slot], 8-byte entries above a sentinel row: WRT_STACK8 below) because any lane may continue any ray.
#pragma once
#include "trace_persistent.cuh"

namespace wrt {

#ifndef WRT_POOL_RAYS
#define WRT_POOL_RAYS 64
#endif
constexpr int kPoolRays = WRT_POOL_RAYS;   // rays per warp (power of two, <= 256)
constexpr int kPoolStack = WRT_STACK_DEPTH;   // stack entries per ray; build_layout refuses trees that could overflow it
#ifndef WRT_POOL_NODE_STEPS
#define WRT_POOL_NODE_STEPS 6
#endif
#ifndef WRT_POOL_PRIM_STEPS
#define WRT_POOL_PRIM_STEPS 4
#endif
#ifndef WRT_POOL_MIN_REFILL
#define WRT_POOL_MIN_REFILL 32
#endif
constexpr int kNodeSteps = WRT_POOL_NODE_STEPS;
constexpr int kPrimSteps = WRT_POOL_PRIM_STEPS;
constexpr unsigned kMinRefill = WRT_POOL_MIN_REFILL;

// __launch_bounds__ second argument of the traversal kernels: 9 blocks of 128 threads caps them at 56 registers, which
// keeps 9 blocks (36 warps, the shared-memory limit) resident per SM; without the cap ptxas takes 67 registers = 7 blocks (C3: 1271 -> 1352 Mrays/s).
#ifndef WRT_MIN_BLOCKS
#define WRT_MIN_BLOCKS 9
#endif
// Tail of a launch: once the global work counter is exhausted and a warp is down to its last rays, those are the rays that
// decide when the launch ends — and on the scenes here they are long because of LEAF work (torus.scene: mean 20, p99.9 1300,
// max 1500 primitive tests per ray).  With at most WRT_COOP_RAYS (16) rays left the warp enters the tail loop, which stops giving one
// lane to one ray: every ray inside a leaf gets a group of 32 / pow2ceil(rays) lanes that test as many records per pass
// (group_leaf_pass below) — 2 - 32 records per ray and pass beat 4 in sequence once the lanes cannot be filled anyway.
#ifndef WRT_COOP_RAYS
#define WRT_COOP_RAYS 16
#endif
#ifndef WRT_COOP_MIN_RECORDS
#define WRT_COOP_MIN_RECORDS 8
#endif
// The NODE work of those last rays is one dependent chain per ray, executed by a warp that has nothing else to issue: what it
// costs is instructions, not memory latency (the longest C3 rays — 2 300 interior visits, 2 100 leaves of ~2 triangles — took
// 3.78 ms alone, 400 ns per visit; prefetching the child pair into L1 changed nothing, profiles/r2_experiments.md).  Most of those
// instructions were the two ring switches per leaf (store the state, re-queue, reload, cooperative pass, re-queue, reload).  In the
// tail loop a node round therefore tests a SHORT leaf (< WRT_COOP_MIN_RECORDS records) right where it meets it — the same
// leaf_record() steps in the same order — and runs WRT_TAIL_NODE_STEPS (48) steps per ray instead of 6 (no other ray waits for the lanes).
#ifndef WRT_MIN_SHARE
#define WRT_MIN_SHARE 8
#endif
#ifndef WRT_TAIL_NODE_STEPS
#define WRT_TAIL_NODE_STEPS 48
#endif
#ifndef WRT_TAIL_INLINE_LEAVES
#define WRT_TAIL_INLINE_LEAVES 1
#endif

// Stack entries in the global scratch.  The reference pushes (far child, t, tmax) and, popping, restores all three
// (KDtreeAccel.cpp:349-384).  tmax is redundant: it only ever changes in the push itself (tmax = t), so at any moment
// T.tmax == the t of the entry below the top (or the root interval's tmax with an empty stack) — by induction over push / pop,
// bit patterns copied, never recomputed.  WRT_STACK8 stores 8-byte entries (node, t) above a sentinel row that holds the root
// tmax; a pop reads rows sp + 1 (node, tmin) and sp (tmax), two independent loads.  Half the L2 footprint and store traffic
// of the 16-byte entries.
#ifndef WRT_STACK8
#define WRT_STACK8 1
#endif

struct PoolSmem {
    float4 a[kPoolRays];               // ox oy oz dx
    float4 b[kPoolRays];               // dy dz ray.tmin ray.tmax
    float4 c[kPoolRays];               // ix iy iz best
    float4 d[kPoolRays];               // T.tmin T.tmax node(int) sp | degenerate << 29 | need_pop << 30 (int)
    int4 e[kPoolRays];                 // res rec rec_end item
    unsigned char ring[3][kPoolRays];  // 0 node, 1 prim, 2 free
};

constexpr size_t kPoolStackBytesPerWarp = (size_t)kPoolStack * kPoolRays * sizeof(float4);

// Cooperative leaf passes of the tail loop (pooled_tail, out of line): with cnt <= 16 rays waiting for primitive tests the warp gives
// every ray a GROUP of G = 32 / pow2ceil(cnt) lanes (one ray: the whole warp).  Lane j of a group tests record rec + j of its ray's
// leaf (Triangle::hit / Sphere::hit with their exact arithmetic; skip records are no-ops: every primitive of the leaf is tested, as
// the reference does), then the hits are offered to the acceptance rule `t - best < -EPS` in RECORD ORDER (KDtreeAccel.cpp:363-373)
// — the same fold the sequential loop performs, so res / best come out identical.  Up to WRT_GROUP_PASSES passes per call.
// Returns, per lane, what became of its group's ray: 0 = still in the leaf, 1 = leaf done (back to the node ring), 2 = decided and
// consumed (boolean queries), 3 = lane has no ray.  The group's leader (j == 0) has written the state back.
#ifndef WRT_GROUP_PASSES
#define WRT_GROUP_PASSES 4
#endif
template <class Src>
__device__ __forceinline__ int group_leaf_pass(const DevSceneView& sc, Src& src, PoolSmem& sm, unsigned slot, bool have, unsigned lane, unsigned lgG)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned G = 1u << lgG, j = lane & (G - 1u), first = lane & ~(G - 1u);
    const unsigned gmask = G == 32u ? FULL : ((1u << G) - 1u);
    const float4 a = sm.a[slot], b = sm.b[slot];
    RayIn r; r.ox = a.x; r.oy = a.y; r.oz = a.z; r.dx = a.w; r.dy = b.x; r.dz = b.y; r.tmin = b.z; r.tmax = b.w;
    const int4 e = sm.e[slot];
    float best = sm.c[slot].w;
    int res = e.x, rec = e.y;
    bool decided = false;
    for (int p = 0; p < WRT_GROUP_PASSES; p++) {
        const int my = rec + (int)j;
        bool hit = false; float t = 0.f; int prim = -1;
        if (have && !decided && my < e.z) {
            const float4* rp = sc.leaf_recs + 3 * (size_t)my;
            const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
            const int kind = __float_as_int(r2.w);
            if (kind == 0) hit = triangle_t(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z, r, t);
            else if (kind == 1) {
                const float lo[3] = { r1.y, r1.z, r1.w }, hi[3] = { r2.x, r2.y, r2.z };
                int inside;
                hit = sphere_t(r0.x, r0.y, r0.z, r1.x, lo, hi, r, t, inside);
            }
            prim = __float_as_int(r0.w);
        }
        unsigned m = (__ballot_sync(FULL, hit) >> first) & gmask;      // this group's hits, bit i = record rec + i
        const float best0 = best;
        while (__any_sync(FULL, m != 0u)) {
            const unsigned i = m ? (unsigned)__ffs(m) - 1u : 0u;
            const float ti = __shfl_sync(FULL, t, first + i);
            const int pi = __shfl_sync(FULL, prim, first + i);
            if (m && (ti - best < -WRT_EPS)) { best = ti; res = pi; }
            m &= m - 1u;
        }
        rec += (int)G;
        if (Src::kCanDecide && have && !decided && best < best0) {
            float tx, ty, tz;
            src.target((size_t)(unsigned)e.w, tx, ty, tz);
            decided = occlusion_decided(r, best, tx, ty, tz);
        }
        if (!__any_sync(FULL, have && !decided && rec < e.z)) break;
    }
    if (!have) return 3;
    const bool leaf_done = rec >= e.z;
    if (j == 0u) {
        if (decided) src.done((size_t)(unsigned)e.w, r, res, best);
        else {
            sm.c[slot].w = best;
            sm.e[slot].x = res; sm.e[slot].y = leaf_done ? e.z : rec;
            if (leaf_done) sm.d[slot].w = __int_as_float(__float_as_int(sm.d[slot].w) | (1 << 30));     // need_pop
        }
    }
    return decided ? 2 : (leaf_done ? 1 : 0);
}


// ---- one ray, the whole warp (tail loop, WRT_PAR_RAYS) ------------------------------------------------------------------------
// What is left of a launch when everything else is done is the dependent chain of its longest rays: ~9 000 visits of ONE ray by a
// warp that has nothing else to issue (C3: 3.2 ms for one grazing camera ray, profiles/r2_experiments.md).  The entries on a ray's
// stack are independent sub-traversals — (node, tmin, tmax) is all the state the reference's loop carries into one — so the warp
// runs them side by side: lane j starts with entry j, every lane walks its sub-tree with a private stack, and a lane that would
// PUSH a far child hands it to an idle lane instead (ballot / rank matching once per step), so the work spreads as fast as the tree
// branches.  What is NOT independent is the reference's acceptance rule `t - best < -EPS` applied in visit order: of two hits closer
// than EPS the one visited first wins.  The lanes therefore do not apply it at all.  Each lane keeps the smallest hit it has seen
// and the second smallest of a different primitive; at the end the warp has the closest hit t1 and the runner-up t2.  If
// t2 - t1 >= 2 EPS the visit order cannot matter: whichever of the two the reference meets first, it ends with t1 (a later t1
// replaces anything >= t1 + EPS, and nothing within EPS of t1 exists to be met first) — that primitive and its t (the same
// Triangle::hit arithmetic, bit for bit) are the answer.  Otherwise (a hit on a shared edge: both neighbours within EPS), or if a
// lane meets the reference's `ray.tmax < tmin` stop (KDtreeAccel.cpp:323, order dependent as well), or a private stack fills up,
// the attempt is dropped: the ray's own state was never touched, it is marked and finishes in the ordinary rounds.
// PRUNED skipping inside a lane uses the lane's own smallest hit, with a 2 EPS guard on rule (a) so that no hit within 2 EPS of the
// final t1 can be skipped anywhere (rule (b) — the ray misses the box — is geometric and needs none).
#ifndef WRT_PAR_RAYS
#define WRT_PAR_RAYS 8
#endif
constexpr int kParDepth = 30;                                    // private stack entries per lane: rows 34..63 of the warp's scratch
constexpr size_t kParBaseF4 = (size_t)34 * kPoolRays * sizeof(float2) / sizeof(float4);
constexpr int kNoParBit = 1 << 28;                               // PoolSmem::d.w: this ray finishes in the ordinary rounds

__device__ __forceinline__ unsigned nth_set_lane(unsigned mask, unsigned k)
{
    for (unsigned i = 0; i < k; i++) mask &= mask - 1u;
    return (unsigned)__ffs(mask) - 1u;
}

struct ParHits { float m1, m2; int p1; };                        // smallest hit (t, primitive) and the second smallest t of another primitive
__device__ __forceinline__ void par_offer(ParHits& H, float t, int prim)
{
    if (prim == H.p1) return;                                    // the same primitive listed by another leaf: the same t
    if (t < H.m1) { H.m2 = H.m1; H.m1 = t; H.p1 = prim; }
    else if (t < H.m2) H.m2 = t;
}

__device__ __forceinline__ bool par_box_prunable(float lox, float loy, float loz, float hix, float hiy, float hiz, const RayIn& r,
                                                 float ix, float iy, float iz, const ParHits& H)
{
    float en, ex;
    box_interval_regular(lox, loy, loz, hix, hiy, hiz, r, ix, iy, iz, en, ex);
    if (H.p1 >= 0 && en > H.m1 * WRT_PRUNE_REL && en > H.m1 + 2.f * WRT_EPS) return true;
    const float m = 1e-4f * (fabsf(en) + fabsf(ex)) + 1e-4f;
    return (en > ex + m) || (ex < -m);
}

// Returns true when the ray was finished (src.done called, the slot is free); false = not decidable this way, state untouched
// except for the kNoParBit mark.  Called by the whole warp, converged; the ray is regular (not degenerate) and waits for a node step.
template <bool PRUNED, class Src>
__device__ __forceinline__ bool par_traverse(const DevSceneView& sc, Src& src, PoolSmem& sm, unsigned slot, float4* __restrict__ gstack, unsigned lane)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const float4 a = sm.a[slot], b = sm.b[slot], c = sm.c[slot], d = sm.d[slot];
    const int4 e = sm.e[slot];
    RayIn r; r.ox = a.x; r.oy = a.y; r.oz = a.z; r.dx = a.w; r.dy = b.x; r.dz = b.y; r.tmin = b.z; r.tmax = b.w;
    const float ix = c.x, iy = c.y, iz = c.z;
    const int spw = __float_as_int(d.w);
    const int sp0 = spw & 0xffff;
    const bool pop0 = (spw >> 30) & 1;
    const float2* col = reinterpret_cast<const float2*>(gstack) + slot;          // the ray's own stack: read only
    float4* priv = gstack + kParBaseF4 + lane;                                     // private entry k at priv[k * 32]
    ParHits H; H.m1 = WRT_INF; H.m2 = WRT_INF; H.p1 = -1;
    if (e.x >= 0) { H.m1 = c.w; H.p1 = e.x; }                                      // the hit the ray already holds (every lane prunes with it)
    // entries in pop order: the current node (unless the ray waits for a pop), then the stack from the top
    const int total = sp0 + (pop0 ? 0 : 1);
    bool active = false, abort = false;
    int node = 0, psp = 0; float tmin = 0.f, tmax = 0.f;
    for (int j = (int)lane; j < total; j += 32) {
        int n_; float t0, t1;
        if (!pop0 && j == 0) { n_ = __float_as_int(d.z); t0 = d.x; t1 = d.y; }
        else {
            const int k = sp0 - (pop0 ? j : j - 1);                                // stack entry k (1-based row), its tmax = t of the row below
            const float2 q = col[(unsigned)k * (unsigned)kPoolRays], below = col[(unsigned)(k - 1) * (unsigned)kPoolRays];
            n_ = __float_as_int(q.x); t0 = q.y; t1 = below.y;
        }
        if (!active) { node = n_; tmin = t0; tmax = t1; active = true; }
        else { priv[(unsigned)psp * 32u] = make_float4(__int_as_float(n_), t0, t1, 0.f); ++psp; }
    }
    for (;;) {
        bool want = false; int far_n = 0; float far_t = 0.f, far_tmax = 0.f;
        bool need_pop = false;
        if (active) {
            if (r.tmax < tmin) abort = true;                                       // KDtreeAccel.cpp:323 ends the WHOLE traversal: order dependent
            const float4* np = sc.nodes + 2 * (size_t)node;
            const float4 na = __ldg(np);
            bool skip = false;
            if (PRUNED) {
                const float4 nb = __ldg(np + 1);
                skip = par_box_prunable(na.z, na.w, nb.x, nb.y, nb.z, nb.w, r, ix, iy, iz, H);
            }
            const unsigned packed = __float_as_uint(na.y);
            const unsigned hi = packed >> 2;
            if (skip) need_pop = true;
            else if ((packed & 3u) == WRT_LEAF_TAG) {
                int rec = __float_as_int(na.x);
                const int end = rec + (int)hi;
                while (rec < end) {
                    const float4* rp = sc.leaf_recs + 3 * (size_t)rec;
                    const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
                    const int kind = __float_as_int(r2.w);
                    rec += 1;
                    if (kind == WRT_REC_SKIP) {
                        if (PRUNED && par_box_prunable(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r, ix, iy, iz, H)) rec += __float_as_int(r0.w);
                        continue;
                    }
                    float t; bool hit;
                    if (kind == 0) hit = triangle_t(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z, r, t);
                    else {
                        const float lo[3] = { r1.y, r1.z, r1.w }, hi3[3] = { r2.x, r2.y, r2.z };
                        int inside;
                        hit = sphere_t(r0.x, r0.y, r0.z, r1.x, lo, hi3, r, t, inside);
                    }
                    if (hit) par_offer(H, t, __float_as_int(r0.w));
                }
                need_pop = true;
            } else {
                const int axis = (int)(packed & 3u);
                const float split = na.x;
                const float o_a = sel3(axis, r.ox, r.oy, r.oz);
                const float d_a = sel3(axis, r.dx, r.dy, r.dz);
                const float i_a = sel3(axis, ix, iy, iz);
                const float t = (split - o_a) * i_a;
                const bool below_first = (o_a < split) || (o_a == split && d_a <= 0.f);
                const int near_n = (int)hi + (below_first ? 0 : 1);
                const int fn = (int)hi + (below_first ? 1 : 0);
                const bool near_only = (t > tmax) || (t <= 0.f);
                const bool far_only = !near_only && (t < tmin);
                if (!near_only && !far_only) { want = true; far_n = fn; far_t = t; far_tmax = tmax; tmax = t; }
                node = far_only ? fn : near_n;
            }
        }
        // lanes that are out of work take over a far child another lane would push
        const bool idle = !active || (need_pop && psp == 0);
        const unsigned pm = __ballot_sync(FULL, want), im = __ballot_sync(FULL, idle);
        const unsigned np_ = (unsigned)__popc(pm), ni = (unsigned)__popc(im);
        const unsigned ri = (unsigned)__popc(im & lt), rp_ = (unsigned)__popc(pm & lt);
        const bool take = idle && ri < np_;
        const unsigned from = take ? nth_set_lane(pm, ri) : 0u;
        const int g_n = __shfl_sync(FULL, far_n, from);
        const float g_t = __shfl_sync(FULL, far_t, from), g_tm = __shfl_sync(FULL, far_tmax, from);
        if (want && rp_ >= ni) {                                                   // nobody free: push on the private stack
            if (psp >= kParDepth) abort = true;
            else { priv[(unsigned)psp * 32u] = make_float4(__int_as_float(far_n), far_t, far_tmax, 0.f); ++psp; }
        }
        if (take) { node = g_n; tmin = g_t; tmax = g_tm; active = true; }
        else if (need_pop || !active) {
            if (psp > 0) { --psp; const float4 q = priv[(unsigned)psp * 32u]; node = __float_as_int(q.x); tmin = q.y; tmax = q.z; active = true; }
            else active = false;
        }
        if (__any_sync(FULL, abort)) break;
        if (!__any_sync(FULL, active)) break;
    }
    bool ok = !__any_sync(FULL, abort);
    int res = -1; float best = WRT_INF;
    if (ok) {
        float g = H.m1;
        for (int o = 16; o > 0; o >>= 1) g = fminf(g, __shfl_sync(FULL, g, (int)(lane ^ (unsigned)o)));
        const unsigned wm = __ballot_sync(FULL, H.p1 >= 0 && H.m1 == g);
        if (wm) {
            res = __shfl_sync(FULL, H.p1, __ffs(wm) - 1);
            best = g;
            float s2 = (H.p1 >= 0 && H.p1 != res) ? H.m1 : H.m2;
            for (int o = 16; o > 0; o >>= 1) s2 = fminf(s2, __shfl_sync(FULL, s2, (int)(lane ^ (unsigned)o)));
            if (s2 - g < 2.f * WRT_EPS) ok = false;                               // two hits the reference's order-dependent rule has to separate
        }
    }
    if (lane == 0) {
        if (ok) src.done((size_t)(unsigned)e.w, r, res, (res >= 0) ? best : WRT_INF);
        else sm.d[slot].w = __int_as_float(spw | kNoParBit);
#ifdef WRT_WARPSIM
        ws_par_count(ok);          // (CPU emulation only: the tests check that this path was taken)
#endif
    }
    return ok;
}

struct PoolRings { unsigned hn, tn, hp, tp, hf, tf; };     // ring heads / tails (monotonic), warp-uniform

// The scheduler loop.  TAIL = false: the steady state described at the top of this file; it returns as soon as the global
// work counter is exhausted and at most WRT_COOP_RAYS rays are left in the pool.  TAIL = true: finishes those last rays —
// longer node rounds with short leaves tested in place, and cooperative leaf passes (group_leaf_pass).  Two instantiations, so
// that the tail's code and registers stay out of the steady-state loop.
template <bool PRUNED, bool TAIL, class Src>
__device__ __forceinline__ void pooled_loop(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n,
                                            PoolSmem& sm, float4* __restrict__ gstack, PoolRings& R)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned RM = kPoolRays - 1;
    unsigned hn = R.hn, tn = R.tn, hp = R.hp, tp = R.tp, hf = R.hf, tf = R.tf;
    bool exhausted = TAIL;
    unsigned share = (unsigned)kPoolRays;
#if !defined(WRT_WARPSIM_FIXED_SHARE)
    if (!TAIL) {
        const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
        const size_t even = (n + warps - 1) / warps;
        if (even < (size_t)kPoolRays) share = even < (size_t)WRT_MIN_SHARE ? (unsigned)WRT_MIN_SHARE : (unsigned)even;
    }
#endif

    for (;;) {
        // ---- refill free slots from the global work counter ------------------------------------
        unsigned nfree = tf - hf;
        // Small launches (fewer work items than pool slots on the device): a warp only fills its even SHARE of the launch
        // (`share`, >= WRT_MIN_SHARE of its 64 slots) and refills when half of that is free, so that the rays of an expensive image
        // region are spread over all resident warps instead of 64 to a warp (torus.scene as shipped: 76 k rays per launch on 5 328 warps).
        if (share < (unsigned)kPoolRays) {
            const unsigned used = (unsigned)kPoolRays - nfree;
            nfree = used + (share + 1u) / 2u <= share ? share - used : 0u;
        }
        if (!TAIL && !exhausted && nfree >= (share < (unsigned)kPoolRays ? 1u : kMinRefill)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)nfree);
            base = __shfl_sync(FULL, base, 0);
            if (base + nfree >= n) exhausted = true;
            for (unsigned k0 = 0; k0 < nfree; k0 += 32) {
                const unsigned k = k0 + lane;
                const bool have = k < nfree;
                unsigned slot = 0;
                if (have) slot = sm.ring[2][(hf + k) & RM];
                __syncwarp();
                bool started = false;
                const size_t item = (size_t)base + k;
                if (have && item < n) {
                    RayIn r;
                    if (src.load(item, r)) {
                        Trav T;
                        if (trav_begin(sc, r, T)) {
                            sm.a[slot] = make_float4(r.ox, r.oy, r.oz, r.dx);
                            sm.b[slot] = make_float4(r.dy, r.dz, r.tmin, r.tmax);
                            sm.c[slot] = make_float4(T.ix, T.iy, T.iz, T.best);
                            sm.d[slot] = make_float4(T.tmin, T.tmax, __int_as_float(0), __int_as_float(T.degen ? (1 << 29) : 0));
                            sm.e[slot] = make_int4(-1, 0, 0, (int)item);
#if WRT_STACK8
                            reinterpret_cast<float2*>(gstack)[slot] = make_float2(0.f, T.tmax);     // sentinel row: root tmax
#endif
                            started = true;
                        } else src.done(item, r, -1, WRT_INF);
                    }
                }
                const unsigned bs = __ballot_sync(FULL, started);
                if (started) sm.ring[0][(tn + __popc(bs & lt)) & RM] = (unsigned char)slot;
                tn += __popc(bs);
                const unsigned bf = __ballot_sync(FULL, have && !started);
                if (have && !started) sm.ring[2][(tf + __popc(bf & lt)) & RM] = (unsigned char)slot;
                tf += __popc(bf);
            }
            hf += nfree;
            __syncwarp();
        }
        const unsigned cn = tn - hn, cp = tp - hp;
        if (cn + cp == 0) {
            if (exhausted) break;
            continue;
        }
        if (!TAIL && WRT_COOP_RAYS > 0 && exhausted && cn + cp <= (unsigned)WRT_COOP_RAYS) break;     // the tail loop takes over
        if (TAIL && WRT_PAR_RAYS > 0 && WRT_STACK8 && cp == 0u && cn <= (unsigned)WRT_PAR_RAYS) {
            // ---- the last rays: one ray, the whole warp (par_traverse) ----------------------------------------------------------
            const unsigned slot = sm.ring[0][hn & RM];
            const int spw = __float_as_int(sm.d[slot].w);
            if (!(spw & kNoParBit) && !((spw >> 29) & 1)) {            // not marked for the ordinary rounds, not degenerate
                hn += 1;
                __syncwarp();
                const bool ok = par_traverse<PRUNED>(sc, src, sm, slot, gstack, lane);
                if (lane == 0) sm.ring[ok ? 2 : 0][(ok ? tf : tn) & RM] = (unsigned char)slot;
                if (ok) tf += 1; else tn += 1;
                __syncwarp();
                continue;
            }
        }
        if (TAIL && cp > 0u && cp >= cn && cp <= 16u) {
            // ---- cooperative leaf passes (tail loop): every ray of the prim ring gets a group of lanes (group_leaf_pass) --------
            const unsigned cnt = cp;
            const unsigned lgG = cnt <= 1u ? 5u : (cnt <= 2u ? 4u : (cnt <= 4u ? 3u : (cnt <= 8u ? 2u : 1u)));
            const unsigned g = lane >> lgG;
            const bool have = g < cnt;
            unsigned slot = 0;
            if (have) slot = sm.ring[1][(hp + g) & RM];
            hp += cnt;
            __syncwarp();
            const int what = group_leaf_pass(sc, src, sm, slot, have, lane, lgG);
            const bool leader = have && (lane & ((1u << lgG) - 1u)) == 0u;
            const unsigned b1 = __ballot_sync(FULL, leader && what == 0);
            if (leader && what == 0) sm.ring[1][(tp + __popc(b1 & lt)) & RM] = (unsigned char)slot;
            tp += __popc(b1);
            const unsigned b0 = __ballot_sync(FULL, leader && what == 1);
            if (leader && what == 1) sm.ring[0][(tn + __popc(b0 & lt)) & RM] = (unsigned char)slot;
            tn += __popc(b0);
            if (Src::kCanDecide) {
                const unsigned b2 = __ballot_sync(FULL, leader && what == 2);
                if (leader && what == 2) sm.ring[2][(tf + __popc(b2 & lt)) & RM] = (unsigned char)slot;
                tf += __popc(b2);
            }
            __syncwarp();
            continue;
        }
        if (cp >= cn) {
            // ---- primitive round ----------------------------------------------------------------
            const unsigned cnt = cp < 32u ? cp : 32u;
            const bool have = lane < cnt;
            unsigned slot = 0;
            if (have) slot = sm.ring[1][(hp + lane) & RM];
            hp += cnt;
            __syncwarp();
            bool leaf_done = false, decided = false;
            if (have) {
                const float4 a = sm.a[slot], b = sm.b[slot];
                RayIn r; r.ox = a.x; r.oy = a.y; r.oz = a.z; r.dx = a.w; r.dy = b.x; r.dz = b.y; r.tmin = b.z; r.tmax = b.w;
                int4 e = sm.e[slot];
                const float4 c = sm.c[slot];
                Trav T; T.ix = c.x; T.iy = c.y; T.iz = c.z; T.best = c.w; T.res = e.x;
                T.degen = (__float_as_int(sm.d[slot].w) >> 29) & 1;
                const float best0 = T.best;
#pragma unroll
                for (int s = 0; s < kPrimSteps; s++) {
                    e.y += leaf_record<PRUNED>(sc, e.y, r, T);
                    if (e.y >= e.z) { leaf_done = true; break; }
                }
                // boolean queries (Scene::occluded) stop once the answer can no longer change (occlusion_decided)
                if (Src::kCanDecide && T.best < best0) {
                    float tx, ty, tz;
                    src.target((size_t)(unsigned)e.w, tx, ty, tz);
                    if (occlusion_decided(r, T.best, tx, ty, tz)) { src.done((size_t)(unsigned)e.w, r, T.res, T.best); decided = true; }
                }
                if (!decided) {
                    if (T.best != best0) { sm.c[slot].w = T.best; }
                    e.x = T.res;
                    sm.e[slot] = e;
                    if (leaf_done) sm.d[slot].w = __int_as_float(__float_as_int(sm.d[slot].w) | (1 << 30));   // need_pop
                }
            }
            const unsigned b1 = __ballot_sync(FULL, have && !leaf_done && !decided);
            if (have && !leaf_done && !decided) sm.ring[1][(tp + __popc(b1 & lt)) & RM] = (unsigned char)slot;
            tp += __popc(b1);
            const unsigned b0 = __ballot_sync(FULL, have && leaf_done && !decided);
            if (have && leaf_done && !decided) sm.ring[0][(tn + __popc(b0 & lt)) & RM] = (unsigned char)slot;
            tn += __popc(b0);
            if (Src::kCanDecide) {
                const unsigned b2 = __ballot_sync(FULL, decided);
                if (decided) sm.ring[2][(tf + __popc(b2 & lt)) & RM] = (unsigned char)slot;
                tf += __popc(b2);
            }
            __syncwarp();
        } else {
            // ---- node round ---------------------------------------------------------------------
            const unsigned cnt = cn < 32u ? cn : 32u;
            const bool have = lane < cnt;
            unsigned slot = 0;
            if (have) slot = sm.ring[0][(hn + lane) & RM];
            hn += cnt;
            __syncwarp();
            int next = 0;     // 0 stays on the node ring, 1 moves to the prim ring, 2 finished
            if (have) {
                const float4 a = sm.a[slot], b = sm.b[slot], c = sm.c[slot], d = sm.d[slot];
                RayIn r; r.ox = a.x; r.oy = a.y; r.oz = a.z; r.dx = a.w; r.dy = b.x; r.dz = b.y; r.tmin = b.z; r.tmax = b.w;
                Trav T; T.ix = c.x; T.iy = c.y; T.iz = c.z; T.best = c.w; T.tmin = d.x; T.tmax = d.y;
                T.node = __float_as_int(d.z);
                const int spw = __float_as_int(d.w);
                T.sp = spw & 0xffff;
                T.degen = (spw >> 29) & 1;
                bool need_pop = (spw >> 30) & 1;
                T.res = sm.e[slot].x;                             // (rec, rec_end, item are only touched when the ray leaves this ring)
                int leaf_first = 0, leaf_end = 0;
#if WRT_STACK8
                float2* stk = reinterpret_cast<float2*>(gstack) + slot;     // entry (sp, slot) at stk[(sp + 1) * kPoolRays]; row 0 = sentinel
#else
                float4* stk = gstack + slot;                      // entry (sp, slot) at stk[sp * kPoolRays]
#endif
                // (measured and dropped: unrolling this loop by 2 -1.4 %, fully -41 % (instruction cache); prefetch.global.L1
                //  of the child pair right after the node header is known -19 % / -42 % for one / both children, of a
                //  leaf's first records on leaf entry -4 %; a warp-uniform loop with a `live` flag instead of the breaks -4 %:
                //  profiles/r1_experiments.md)
                for (int s = 0; s < (TAIL ? WRT_TAIL_NODE_STEPS : kNodeSteps); s++) {
                    if (need_pop) {
                        need_pop = false;
                        if (T.sp <= 0) { next = 2; break; }
                        --T.sp;
#if WRT_STACK8
                        const float2 q = stk[(unsigned)(T.sp + 1) * (unsigned)kPoolRays];
                        const float2 below = stk[(unsigned)T.sp * (unsigned)kPoolRays];
                        T.node = __float_as_int(q.x); T.tmin = q.y; T.tmax = below.y;
#else
                        const float4 q = stk[(unsigned)T.sp * (unsigned)kPoolRays];
                        T.node = __float_as_int(q.x); T.tmin = q.y; T.tmax = q.z;
#endif
                    }
                    if (r.tmax < T.tmin) { next = 2; break; }                              // KDtreeAccel.cpp:323
                    const float4* np = sc.nodes + 2 * (size_t)T.node;
                    const float4 na = __ldg(np);
                    bool skip = false;
                    if (PRUNED) {
                        const float4 nb = __ldg(np + 1);
                        skip = box_prunable(na.z, na.w, nb.x, nb.y, nb.z, nb.w, r, T.ix, T.iy, T.iz, T.res, T.best, T.degen);
                    }
                    const unsigned packed = __float_as_uint(na.y);
                    const unsigned hi = packed >> 2;                  // child pair (interior) | record count (leaf)
                    const bool leaf = (packed & 3u) == WRT_LEAF_TAG;
                    if (skip || (leaf && hi == 0u)) { need_pop = true; continue; }
                    if (leaf) {
                        if (TAIL && WRT_TAIL_INLINE_LEAVES && hi < (unsigned)WRT_COOP_MIN_RECORDS) {      // short leaf, tail loop: tested here
                            int rec = __float_as_int(na.x);
                            const int end = rec + (int)hi;
                            const float best0 = T.best;
                            while (rec < end) rec += leaf_record<PRUNED>(sc, rec, r, T);
                            if (Src::kCanDecide && T.best < best0) {
                                float tx, ty, tz;
                                src.target((size_t)(unsigned)sm.e[slot].w, tx, ty, tz);
                                if (occlusion_decided(r, T.best, tx, ty, tz)) { next = 2; break; }
                            }
                            need_pop = true;
                            continue;
                        }
                        leaf_first = __float_as_int(na.x); leaf_end = leaf_first + (int)hi; next = 1; break;
                    }
                    // interior step (trav_interior, KDtreeAccel.cpp:325-358) without branches; the push goes to the
                    // global scratch stack
                    const int axis = (int)(packed & 3u);
                    const float split = na.x;
                    const float o_a = sel3(axis, r.ox, r.oy, r.oz);
                    const float d_a = sel3(axis, r.dx, r.dy, r.dz);
                    const float i_a = sel3(axis, T.ix, T.iy, T.iz);
                    const float t = (split - o_a) * i_a;
                    const bool below_first = (o_a < split) || (o_a == split && d_a <= 0.f);
                    const int near_n = (int)hi + (below_first ? 0 : 1);
                    const int far_n = (int)hi + (below_first ? 1 : 0);
                    const bool near_only = (t > T.tmax) || (t <= 0.f);
                    const bool far_only = !near_only && (t < T.tmin);
                    const bool both = !near_only && !far_only;
#if WRT_STACK8
                    if (both && T.sp < kPoolStack) { stk[(unsigned)(T.sp + 1) * (unsigned)kPoolRays] = make_float2(__int_as_float(far_n), t); ++T.sp; }
#else
                    if (both && T.sp < kPoolStack) { stk[(unsigned)T.sp * (unsigned)kPoolRays] = make_float4(__int_as_float(far_n), t, T.tmax, 0.f); ++T.sp; }
#endif
                    T.node = far_only ? far_n : near_n;
                    T.tmax = both ? t : T.tmax;
                }
                if (next == 2) {
                    src.done((size_t)(unsigned)sm.e[slot].w, r, T.res, (T.res >= 0) ? T.best : WRT_INF);
                } else {
                    sm.d[slot] = make_float4(T.tmin, T.tmax, __int_as_float(T.node), __int_as_float(T.sp | (T.degen ? (1 << 29) : 0) | (need_pop ? (1 << 30) : 0) | (spw & kNoParBit)));
                    if (next == 1) { sm.e[slot].y = leaf_first; sm.e[slot].z = leaf_end; }
                    if (TAIL && WRT_TAIL_INLINE_LEAVES) { sm.c[slot].w = T.best; sm.e[slot].x = T.res; }      // (leaves tested in this round)
                }
            }
            const unsigned b0 = __ballot_sync(FULL, have && next == 0);
            if (have && next == 0) sm.ring[0][(tn + __popc(b0 & lt)) & RM] = (unsigned char)slot;
            tn += __popc(b0);
            const unsigned b1 = __ballot_sync(FULL, have && next == 1);
            if (have && next == 1) sm.ring[1][(tp + __popc(b1 & lt)) & RM] = (unsigned char)slot;
            tp += __popc(b1);
            const unsigned b2 = __ballot_sync(FULL, have && next == 2);
            if (have && next == 2) sm.ring[2][(tf + __popc(b2 & lt)) & RM] = (unsigned char)slot;
            tf += __popc(b2);
            __syncwarp();
        }
    }
    R.hn = hn; R.tn = tn; R.hp = hp; R.tp = tp; R.hf = hf; R.tf = tf;
}

template <bool PRUNED, class Src>
#if !defined(WRT_WARPSIM)
__device__ __noinline__
#endif
void pooled_tail(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n, PoolSmem& sm, float4* __restrict__ gstack, PoolRings& R)
{
    pooled_loop<PRUNED, true>(sc, src, counter, n, sm, gstack, R);
}

template <bool PRUNED, class Src>
__device__ __forceinline__ void trace_pooled(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n,
                                             PoolSmem& sm, float4* __restrict__ gstack)
{
    const unsigned lane = threadIdx.x & 31;
    PoolRings R = { 0u, 0u, 0u, 0u, 0u, (unsigned)kPoolRays };
    for (unsigned i = lane; i < (unsigned)kPoolRays; i += 32) sm.ring[2][i] = (unsigned char)i;
    __syncwarp();
    pooled_loop<PRUNED, false>(sc, src, counter, n, sm, gstack, R);
    if ((R.tn - R.hn) + (R.tp - R.hp) != 0u) pooled_tail<PRUNED>(sc, src, counter, n, sm, gstack, R);
}

// Compile-time choice of the scheduler used by the kernels (A/B measured in profiles/):
// 2 = lane refill + vote, 3 = pooled (default).  Measured and removed (profiles/r1_experiments.md, "Schedulers 4 and 5"):
// scheduler 1 (lane refill + while-while), scheduler 4 (pooled + child-pair node visits: both children fetched and
// bounds-tested at the parent) and scheduler 5 (k rays owned by every lane, state in conflict-free shared memory).
#ifndef WRT_TRACE_SCHED
#define WRT_TRACE_SCHED 3
#endif
// Scheduler 0: each warp pulls 32 consecutive work items and every lane runs kd_traverse() to completion.
// No refill, no votes: cheapest bookkeeping, right for trees of a few dozen nodes (C4's Cornell box).
template <bool PRUNED, class Src>
__device__ __forceinline__ void trace_plain(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n)
{
    for (;;) {
        unsigned long long b = 0;
        if ((threadIdx.x & 31) == 0) b = atomicAdd(counter, 32ull);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= n) break;
        const size_t i = (size_t)b + (threadIdx.x & 31);
        if (i >= n) continue;
        RayIn r;
        if (!src.load(i, r)) continue;
        float t;
        const int prim = kd_traverse<PRUNED, false>(sc, r, t, nullptr);
        src.done(i, r, prim, t);
    }
}

template <bool PRUNED, class Src>
__device__ __forceinline__ void trace_rays(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n,
                                           float4* trav_scratch)
{
    if (sc.small_tree) { trace_plain<PRUNED>(sc, src, counter, n); return; }
#if WRT_TRACE_SCHED == 3
    __shared__ PoolSmem pool_smem[4];                       // kernels launch 128 threads = 4 warps
    const unsigned warp = threadIdx.x >> 5;
    float4* stk = trav_scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * ((size_t)kPoolStack * kPoolRays);
    trace_pooled<PRUNED>(sc, src, counter, n, pool_smem[warp], stk);
#else
    trace_persistent_vote<PRUNED>(sc, src, counter, n);
#endif
}

// Global scratch for the pooled scheduler's traversal stacks: one region per resident warp of a launch.
inline size_t trav_scratch_bytes(int grid_blocks, int block_threads)
{
    return (size_t)grid_blocks * (size_t)(block_threads / 32) * kPoolStackBytesPerWarp;
}

}  // namespace wrt
