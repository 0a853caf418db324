// Persistent-warp ray traversal with lane refill.
//
// One thread = one ray, but a warp does not wait for its slowest ray: whenever fewer than
// kRefillThreshold lanes are still traversing, the idle lanes claim the next work items from a global
// counter (one atomicAdd per refill, ballot/popc ranks) and start new rays while the busy lanes keep
// theirs.  Rays in this tree differ by two orders of magnitude in length (SIMT efficiency without
// refill: ~3 of 32 lanes, profiles/r1_ncu_extend_v0_summary.md).
// Inside the work loop every busy lane is either at a node or inside a leaf, and each round the warp
// votes which of the two code paths to run (trace_persistent_vote below).
//
// `Src` supplies the work:  bool load(size_t item, RayIn& r)  — build the ray of work item `item`
//                           void done(size_t item, const RayIn& r, int prim, float t) — consume the result
// The per-ray steps are the ones of traverse.cuh, so results are identical to kd_traverse().
#pragma once
#include "traverse.cuh"

namespace wrt {

#ifndef WRT_REFILL_THRESHOLD
#define WRT_REFILL_THRESHOLD 28
#endif
constexpr int kRefillThreshold = WRT_REFILL_THRESHOLD;

#ifndef WRT_FAST_DIV
#define WRT_FAST_DIV 1
#endif

// Triangle::hit + the acceptance rule `t - best < -EPS`, answering only "does this triangle become the
// new best, and with which t".  The reference only COMPARES beta, gamma and beta+gamma against
// thresholds, and t against EPS, [tmin,tmax] and best-EPS; only a winning t is kept.  So the three IEEE
// divisions are replaced by one reciprocal approximation (MUFU.RCP, ~1 ulp) whenever the quotient is far
// from every threshold it is compared with (margin 1e-5 relative, >= 40x the approximation error);
// numerators and the denominator are still the reference's exact no-FMA expressions.  Anything within
// the margin, any NaN/inf and any extreme denominator takes the exact path (triangle_t), and a winning t
// is always produced by the IEEE division, so results stay bit-identical (GPU tests compare >1e7 rays).
__device__ __forceinline__ bool triangle_wins(float p0x, float p0y, float p0z, float A, float B, float C,
                                              float D, float E, float F, const RayIn& r, float best, float& t_out)
{
#if WRT_FAST_DIV
    const float G = r.dx, H = r.dy, I = r.dz;
    const float J = p0x - r.ox, K = p0y - r.oy, L = p0z - r.oz;
    const float EIHF = E * I - H * F;
    const float GFDI = G * F - D * I;
    const float DHEG = D * H - E * G;
    const float denom = A * EIHF + B * GFDI + C * DHEG;
    const float ad = fabsf(denom);
    if (ad > 1e-30f && ad < 1e30f) {
        float rd;
#if defined(WRT_WARPSIM)
        rd = 1.0f / denom;      // tests/hostsim/warpsim.cpp (CPU emulation of a warp): any reciprocal inside the margins decides alike
#else
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rd) : "f"(denom));
#endif
        const float bq = (J * EIHF + K * GFDI + L * DHEG) * rd;
        const float mb = 1e-5f * fabsf(bq) + 2e-8f;
        if (bq < -WRT_EPS - mb || bq > 1.f + mb) return false;                       // certainly rejected on beta
        if (bq > -WRT_EPS + mb && bq < 1.f - mb) {                                   // certainly passes beta
            const float AKJB = A * K - J * B;
            const float JCAL = J * C - A * L;
            const float BLKC = B * L - K * C;
            const float gq = (I * AKJB + H * JCAL + G * BLKC) * rd;
            const float mg = 1e-5f * fabsf(gq) + 2e-8f;
            const float sum = bq + gq;
            if (gq < -WRT_EPS - mg || sum > 1.f + 2e-5f) return false;               // certainly rejected on gamma
            if (gq > -WRT_EPS + mg && sum < 1.f - 2e-5f) {                           // certainly inside
                const float nt = -(F * AKJB + E * JCAL + D * BLKC);
                const float tq = nt * rd;
                const float mt = 1e-5f * fabsf(tq) + 1e-8f;
                const float thr = best - WRT_EPS;
                const float m2 = 1e-5f * (fabsf(tq) + fabsf(best)) + 1e-7f;
                if (tq < WRT_EPS - mt || tq < r.tmin - mt || tq > r.tmax + mt || tq > thr + m2) return false;   // cannot win
                if (tq > WRT_EPS + mt && tq > r.tmin + mt && tq < r.tmax - mt && tq < thr - m2) {
                    t_out = nt / denom;                                              // the reference's t, IEEE division
                    return true;
                }
            }
        }
    }
#endif
    float t;
    if (!triangle_t(p0x, p0y, p0z, A, B, C, D, E, F, r, t)) return false;
    if (!(t - best < -WRT_EPS)) return false;
    t_out = t;
    return true;
}

// Scene::occluded / shadowRayTest only ask whether the closest hit point equals the target within EPS per
// component (scene.cpp:64-67).  `best` only decreases during a traversal, so once the current best hit lies
// BEFORE the target along some axis by clearly more than EPS, every later (smaller) best does too.
__device__ __forceinline__ bool occlusion_decided(const RayIn& r, float best, float tx, float ty, float tz)
{
    const float m = 1.5f * WRT_EPS;
    const float ex = (r.ox + r.dx * best) - tx, ey = (r.oy + r.dy * best) - ty, ez = (r.oz + r.dz * best) - tz;
    return (r.dx > 0.f ? ex < -m : (r.dx < 0.f && ex > m)) || (r.dy > 0.f ? ey < -m : (r.dy < 0.f && ey > m)) ||
           (r.dz > 0.f ? ez < -m : (r.dz < 0.f && ez > m));
}

// One leaf record against the ray + the reference's acceptance rule (KDtreeAccel.cpp:363-373).
// Returns how many records to advance: 1, or 1 + n for a prunable skip record (PRUNED only).
// (TT: any traversal state with `best`, `res`, invDir `ix, iy, iz` and `degen`.)
template <bool PRUNED, class TT>
__device__ __forceinline__ int leaf_record_regs(const float4 r0, const float4 r1, const float4 r2, const RayIn& r, TT& T)
{
    if (__float_as_int(r2.w) == WRT_REC_SKIP) {
        if (PRUNED && box_prunable(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r, T.ix, T.iy, T.iz, T.res, T.best, T.degen)) return 1 + __float_as_int(r0.w);
        return 1;
    }
    if (__float_as_int(r2.w) == 0) {
        float t;
        if (triangle_wins(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, r2.x, r2.y, r2.z, r, T.best, t)) { T.best = t; T.res = __float_as_int(r0.w); }
    } else {
        const float lo[3] = { r1.y, r1.z, r1.w }, hi[3] = { r2.x, r2.y, r2.z };
        int inside; float t;
        if (sphere_t(r0.x, r0.y, r0.z, r1.x, lo, hi, r, t, inside) && (t - T.best < -WRT_EPS)) { T.best = t; T.res = __float_as_int(r0.w); }
    }
    return 1;
}

template <bool PRUNED, class TT>
__device__ __forceinline__ int leaf_record(const DevSceneView& sc, int rec_index, const RayIn& r, TT& T)
{
    const float4* rec = sc.leaf_recs + 3 * (size_t)rec_index;
    const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
    return leaf_record_regs<PRUNED>(r0, r1, r2, r, T);
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
    bool active = false, in_leaf = false, need_pop = false;
    int rec = 0, rec_end = 0;
    bool exhausted = false;
    for (;;) {
        if (!exhausted) {
            const unsigned need = __ballot_sync(FULL, !active);
            if (need) {
                const int cnt = __popc(need);
                const int leader = __ffs(need) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)cnt);
                base = __shfl_sync(FULL, base, leader);
                if (base + (unsigned long long)cnt >= n) exhausted = true;
                if (!active) {
                    const size_t idx = (size_t)base + __popc(need & ((1u << lane) - 1u));
                    if (idx < n) {
                        item = idx;
                        if (src.load(idx, r)) {
                            if (trav_begin(sc, r, T)) { active = true; in_leaf = false; need_pop = false; }
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
            if (2 * n_leaf >= n_busy) {
                // primitive round: up to two leaf records per lane that is inside a leaf.  A lane that
                // exhausts its leaf only flags `need_pop`; the pop itself runs in the next node round,
                // where most lanes do it together.
                if (active && in_leaf) {
                    // (loading both records up front was measured: 79 registers, 346 -> 281 Mrays/s)
                    const float best0 = T.best;
                    rec += leaf_record<PRUNED>(sc, rec, r, T);
                    if (rec >= rec_end) { in_leaf = false; need_pop = true; }
                    else {
                        rec += leaf_record<PRUNED>(sc, rec, r, T);
                        if (rec >= rec_end) { in_leaf = false; need_pop = true; }
                    }
                    // boolean queries (Scene::occluded) may stop once the answer can no longer change
                    if (T.best < best0 && src.decided(r, T.best)) {
                        src.done(item, r, T.res, T.best);
                        active = false; in_leaf = false; need_pop = false;
                    }
                }
            } else {
                // node round: (pop, then) one node visit per lane that is at a node
                if (active && !in_leaf) {
                    bool finished = false;
                    if (need_pop) { need_pop = false; finished = !trav_pop(T, S); }
                    if (!finished) {
                        if (r.tmax < T.tmin) finished = true;                      // KDtreeAccel.cpp:323
                        else {
                            const float4 na = __ldg(&sc.nodes[2 * T.node]);
                            if (trav_skip<PRUNED>(sc, na, r, T)) need_pop = true;
                            else if ((__float_as_uint(na.y) & 3u) == WRT_LEAF_TAG) {
                                const int cnt = (int)(__float_as_uint(na.y) >> 2);
                                if (cnt > 0) { in_leaf = true; rec = __float_as_int(na.x); rec_end = rec + cnt; }
                                else need_pop = true;
                            } else trav_interior(na, r, T, S);
                        }
                    }
                    if (finished) { src.done(item, r, T.res, (T.res >= 0) ? T.best : WRT_INF); active = false; }
                }
            }
        }
    }
}

}  // namespace wrt
