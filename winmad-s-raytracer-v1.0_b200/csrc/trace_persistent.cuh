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

}  // namespace wrt
