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
// scratch area ([depth][slot], 16-byte entries) because any lane may continue any ray.
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
#ifndef WRT_TOS_CACHE
#define WRT_TOS_CACHE 1
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

template <bool PRUNED, class Src>
__device__ __forceinline__ void trace_pooled(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n,
                                             PoolSmem& sm, float4* __restrict__ gstack)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned RM = kPoolRays - 1;
    unsigned hn = 0, tn = 0, hp = 0, tp = 0, hf = 0, tf = kPoolRays;   // ring heads / tails (monotonic)
    for (unsigned i = lane; i < (unsigned)kPoolRays; i += 32) sm.ring[2][i] = (unsigned char)i;
    __syncwarp();
    bool exhausted = false;

    for (;;) {
        // ---- refill free slots from the global work counter ------------------------------------
        unsigned nfree = tf - hf;
        if (!exhausted && nfree >= kMinRefill) {
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
                float4* stk = gstack + slot;                      // entry (sp, slot) at stk[sp * kPoolRays]
#if WRT_TOS_CACHE
                // The newest stack entry stays in registers while the lane works on the ray: most pops follow their push within
                // a step or two (near child skipped by its bounds / empty leaf), and then neither the store nor the dependent
                // L2 load of the scratch stack happens.  T.sp counts the entries in the scratch stack only.
                int tos_node = -1; float tos_t = 0.f, tos_tmax = 0.f;
#endif
                // (measured and dropped: unrolling this loop by 2 -1.4 %, fully -41 % (instruction cache); prefetch.global.L1
                //  of the child pair right after the node header is known -19 % / -42 % for one / both children, of a
                //  leaf's first records on leaf entry -4 %; a warp-uniform loop with a `live` flag instead of the breaks -4 %:
                //  profiles/r1_experiments.md)
                for (int s = 0; s < kNodeSteps; s++) {
                    if (need_pop) {
                        need_pop = false;
#if WRT_TOS_CACHE
                        if (tos_node >= 0) { T.node = tos_node; T.tmin = tos_t; T.tmax = tos_tmax; tos_node = -1; }
                        else
#endif
                        {
                            if (T.sp <= 0) { next = 2; break; }
                            --T.sp;
                            const float4 q = stk[(unsigned)T.sp * (unsigned)kPoolRays];
                            T.node = __float_as_int(q.x); T.tmin = q.y; T.tmax = q.z;
                        }
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
                    if (leaf) { leaf_first = __float_as_int(na.x); leaf_end = leaf_first + (int)hi; next = 1; break; }
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
#if WRT_TOS_CACHE
                    if (both) {
                        if (tos_node >= 0 && T.sp < kPoolStack) { stk[(unsigned)T.sp * (unsigned)kPoolRays] = make_float4(__int_as_float(tos_node), tos_t, tos_tmax, 0.f); ++T.sp; }
                        tos_node = far_n; tos_t = t; tos_tmax = T.tmax;
                    }
#else
                    if (both && T.sp < kPoolStack) { stk[(unsigned)T.sp * (unsigned)kPoolRays] = make_float4(__int_as_float(far_n), t, T.tmax, 0.f); ++T.sp; }
#endif
                    T.node = far_only ? far_n : near_n;
                    T.tmax = both ? t : T.tmax;
                }
                if (next == 2) {
                    src.done((size_t)(unsigned)sm.e[slot].w, r, T.res, (T.res >= 0) ? T.best : WRT_INF);
                } else {
#if WRT_TOS_CACHE
                    if (tos_node >= 0 && T.sp < kPoolStack) { stk[(unsigned)T.sp * (unsigned)kPoolRays] = make_float4(__int_as_float(tos_node), tos_t, tos_tmax, 0.f); ++T.sp; }
#endif
                    sm.d[slot] = make_float4(T.tmin, T.tmax, __int_as_float(T.node), __int_as_float(T.sp | (T.degen ? (1 << 29) : 0) | (need_pop ? (1 << 30) : 0)));
                    if (next == 1) { sm.e[slot].y = leaf_first; sm.e[slot].z = leaf_end; }
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
