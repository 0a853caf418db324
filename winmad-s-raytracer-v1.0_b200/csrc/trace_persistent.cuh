// Persistent-warp ray traversal with lane refill.
//
// One thread = one ray, but a warp does not wait for its slowest ray: whenever fewer than
// kRefillThreshold lanes are still traversing, the idle lanes claim the next work items from a global
// counter (one atomicAdd per refill, ballot/popc ranks) and start new rays while the busy lanes keep
// theirs.  Rays in this tree differ by two orders of magnitude in length (SIMT efficiency without
// refill: ~3 of 32 lanes, profiles/r1_ncu_extend_summary.md), so this is the main lever.
// Inside the work loop the warp alternates between "every busy lane walks down to its next leaf" and
// "every busy lane intersects its leaf" (while-while), so both phases run converged.
//
// `Src` supplies the work:  bool load(size_t item, RayIn& r)  — build the ray of work item `item`
//                           void done(size_t item, const RayIn& r, int prim, float t) — consume the result
// The per-ray steps are the ones of traverse.cuh, so results are identical to kd_traverse().
#pragma once
#include "traverse.cuh"

namespace wrt {

constexpr int kRefillThreshold = 22;

// One leaf record against the ray + the reference's acceptance rule (KDtreeAccel.cpp:363-373).
__device__ __forceinline__ void leaf_record(const DevSceneView& sc, int rec_index, const RayIn& r, Trav& T)
{
    const float4* rec = sc.leaf_recs + 3 * (size_t)rec_index;
    const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
    float t; bool hit;
    if (__float_as_int(r2.w) == 0) {
        hit = triangle_t(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z, r, t);
    } else {
        const float lo[3] = { r1.y, r1.z, r1.w }, hi[3] = { r2.x, r2.y, r2.z };
        int inside;
        hit = sphere_t(r0.x, r0.y, r0.z, r1.x, lo, hi, r, t, inside);
    }
    if (hit && (t - T.best < -WRT_EPS)) { T.best = t; T.res = __float_as_int(r0.w); }
}

// Scheduler 2 ("vote"): every busy lane is either at a node or inside a leaf.  Each round the warp
// votes and runs ONE of the two code paths — a node visit or one primitive test — for the lanes that
// want it; the others wait a round.  Both paths therefore execute with most of the busy lanes on
// board, instead of the while-while form where a warp waits for its deepest descent and its largest leaf.
template <bool PRUNED, class Src>
__device__ __forceinline__ void trace_persistent_vote(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    Trav T;
    TravStack S;
    RayIn r;
    size_t item = 0;
    bool active = false, in_leaf = false;
    int rec = 0, rec_end = 0;
    bool exhausted = false;
    for (;;) {
        if (!exhausted) {
            const unsigned need = __ballot_sync(FULL, !active);
            if (need) {
                const int cnt = __popc(need);
                const int leader = __ffs(need) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(counter, (unsigned long long)cnt);
                base = __shfl_sync(FULL, base, leader);
                if (base + (unsigned long long)cnt >= n) exhausted = true;
                if (!active) {
                    const size_t idx = (size_t)base + __popc(need & ((1u << lane) - 1u));
                    if (idx < n) {
                        item = idx;
                        if (src.load(idx, r)) {
                            if (trav_begin(sc, r, T)) { active = true; in_leaf = false; }
                            else src.done(idx, r, -1, WRT_INF);
                        }
                    }
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }
        for (;;) {
            const unsigned m_busy = __ballot_sync(FULL, active);
            const unsigned m_leaf = __ballot_sync(FULL, active && in_leaf);
            const int n_busy = __popc(m_busy), n_leaf = __popc(m_leaf);
            if (n_busy == 0 || (!exhausted && n_busy < kRefillThreshold)) break;
            if (3 * n_leaf >= n_busy) {
                // primitive round: one leaf record per lane that is inside a leaf
                if (active && in_leaf) {
                    leaf_record(sc, rec, r, T);
                    if (++rec == rec_end) {
                        in_leaf = false;
                        if (!trav_pop(T, S)) { src.done(item, r, T.res, (T.res >= 0) ? T.best : WRT_INF); active = false; }
                    }
                }
            } else {
                // node round: one node visit per lane that is at a node
                if (active && !in_leaf) {
                    bool finished = false;
                    if (r.tmax < T.tmin) finished = true;                          // KDtreeAccel.cpp:323
                    else {
                        const float4 na = __ldg(&sc.nodes[2 * T.node]);
                        if (trav_skip<PRUNED>(sc, na, r, T)) finished = !trav_pop(T, S);
                        else if ((__float_as_uint(na.y) & 3u) == WRT_LEAF_TAG) {
                            const int cnt = (int)(__float_as_uint(na.y) >> 2);
                            if (cnt > 0) { in_leaf = true; rec = __float_as_int(na.x); rec_end = rec + cnt; }
                            else finished = !trav_pop(T, S);
                        } else trav_interior(na, r, T, S);
                    }
                    if (finished) { src.done(item, r, T.res, (T.res >= 0) ? T.best : WRT_INF); active = false; }
                }
            }
        }
    }
}


template <bool PRUNED, class Src>
__device__ __forceinline__ void trace_persistent(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    Trav T;
    TravStack S;
    RayIn r;
    size_t item = 0;
    bool active = false;
    bool exhausted = false;     // warp-uniform: the counter ran past n
    for (;;) {
        if (!exhausted) {
            const unsigned need = __ballot_sync(FULL, !active);
            if (need) {
                const int cnt = __popc(need);
                const int leader = __ffs(need) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(counter, (unsigned long long)cnt);
                base = __shfl_sync(FULL, base, leader);
                if (base + (unsigned long long)cnt >= n) exhausted = true;
                if (!active) {
                    const size_t idx = (size_t)base + __popc(need & ((1u << lane) - 1u));
                    if (idx < n) {
                        item = idx;
                        if (src.load(idx, r)) {
                            if (trav_begin(sc, r, T)) active = true;
                            else src.done(idx, r, -1, WRT_INF);
                        }
                    }
                }
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }
        for (;;) {
            // phase 1: every busy lane walks down to its next leaf (or finishes)
            bool at_leaf = false;
            float4 na = make_float4(0.f, 0.f, 0.f, 0.f);
            while (active) {
                if (r.tmax < T.tmin) {                                            // KDtreeAccel.cpp:323
                    src.done(item, r, T.res, (T.res >= 0) ? T.best : WRT_INF);
                    active = false;
                    break;
                }
                na = __ldg(&sc.nodes[2 * T.node]);
                if (trav_skip<PRUNED>(sc, na, r, T)) {
                    if (!trav_pop(T, S)) {
                        src.done(item, r, T.res, (T.res >= 0) ? T.best : WRT_INF);
                        active = false;
                    }
                    continue;
                }
                if ((__float_as_uint(na.y) & 3u) == WRT_LEAF_TAG) { at_leaf = true; break; }
                trav_interior(na, r, T, S);
            }
            // phase 2: every busy lane intersects its leaf, then pops
            if (active && at_leaf) {
                trav_leaf<false>(sc, na, r, T, nullptr);
                if (!trav_pop(T, S)) {
                    src.done(item, r, T.res, (T.res >= 0) ? T.best : WRT_INF);
                    active = false;
                }
            }
            const int busy = __popc(__ballot_sync(FULL, active));
            if (busy == 0 || (!exhausted && busy < kRefillThreshold)) break;
        }
    }
}

// Compile-time choice of the scheduler used by the kernels (A/B measured in profiles/).
#ifndef WRT_TRACE_SCHED
#define WRT_TRACE_SCHED 2
#endif
template <bool PRUNED, class Src>
__device__ __forceinline__ void trace_rays(const DevSceneView& sc, Src& src, unsigned long long* counter, size_t n)
{
#if WRT_TRACE_SCHED == 2
    trace_persistent_vote<PRUNED>(sc, src, counter, n);
#else
    trace_persistent<PRUNED>(sc, src, counter, n);
#endif
}

}  // namespace wrt
