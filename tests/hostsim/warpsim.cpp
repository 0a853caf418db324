// TEST-ONLY library ("warpsim"): runs the product's WARP-LEVEL traversal schedulers — trace_pooled (scheduler 3, the
// default) and trace_persistent_vote (scheduler 2) from csrc/trace_pooled.cuh / trace_persistent.cuh — on the CPU.  One
// std::thread per lane (128 per block = 4 warps, as the kernels launch), CUDA's warp primitives (__ballot_sync, __shfl_sync,
// __any_sync, __syncwarp, atomicAdd) emulated with barriers, `__shared__` arrays as statics.  The headers are compiled
// unmodified (-DWRT_WARPSIM only swaps the rcp.approx inline PTX for a division), so the ring / refill / re-queue logic that
// `pytest -m gpu` exercises on a B200 can also be checked against the oracle where there is no GPU.  Never part of the product.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// ---- CUDA emulation layer (must precede the product headers) ------------------------------------------------------------
struct WsDim3 { unsigned x, y, z; };
static thread_local WsDim3 threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 };
static WsDim3 blockDim = { 128, 1, 1 }, gridDim = { 1, 1, 1 };
#define __device__
#define __global__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)

class WsBarrier {
public:
    explicit WsBarrier(int n) : n_(n), waiting_(0), gen_(0) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned g = gen_;
        if (++waiting_ == n_) { waiting_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != g; });
    }
private:
    std::mutex m_; std::condition_variable cv_; int n_, waiting_; unsigned gen_;
};

struct WsWarp { WsBarrier bar{ 32 }; unsigned long long slot[32]; };
static thread_local WsWarp* ws_warp = nullptr;
static thread_local unsigned ws_lane = 0;
static WsBarrier* ws_block_bar = nullptr;

static inline unsigned __ballot_sync(unsigned, bool pred)
{
    ws_warp->slot[ws_lane] = pred ? 1ull : 0ull;
    ws_warp->bar.wait();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) if (ws_warp->slot[i]) m |= 1u << i;
    ws_warp->bar.wait();
    return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long w = 0; memcpy(&w, &v, sizeof v);
    ws_warp->slot[ws_lane] = w;
    ws_warp->bar.wait();
    T r; memcpy(&r, &ws_warp->slot[src & 31], sizeof r);
    ws_warp->bar.wait();
    return r;
}
static inline void __syncwarp() { ws_warp->bar.wait(); }
static inline void __syncthreads() { ws_block_bar->wait(); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
struct alignas(16) int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = { x, y, z, w }; return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }

static std::atomic<unsigned long long> g_par_ok{0}, g_par_fallback{0};
static inline void ws_par_count(bool ok) { if (ok) g_par_ok++; else g_par_fallback++; }
#include "scene_layout.h"
#include "trace_pooled.cuh"      // includes trace_persistent.cuh and traverse.cuh

using namespace wrt;
namespace wrt { void set_error(const std::string&) {} }

// ---- work sources (the kernels' ClosestSrc / OccludedSrc, host flavour) ------------------------------------------------
struct WsClosestSrc {
    const wrt_ray* rays; int32_t* prim; float* t_out;
    static constexpr bool kCanDecide = false;
    void target(size_t, float&, float&, float&) const {}
    bool decided(const RayIn&, float) const { return false; }
    bool load(size_t i, RayIn& r) const
    {
        const wrt_ray& q = rays[i];
        r.ox = q.ox; r.oy = q.oy; r.oz = q.oz; r.dx = q.dx; r.dy = q.dy; r.dz = q.dz; r.tmin = q.tmin; r.tmax = q.tmax;
        return true;
    }
    void done(size_t i, const RayIn&, int id, float t) const { prim[i] = id; t_out[i] = t; }
};

struct WsOccludedSrc {
    const float* q9; uint8_t* occluded;
    float tx, ty, tz;      // scheduler 2 keeps the target of the lane's ray here
    static constexpr bool kCanDecide = true;
    void target(size_t i, float& x, float& y, float& z) const { x = q9[9 * i + 6]; y = q9[9 * i + 7]; z = q9[9 * i + 8]; }
    bool load(size_t i, RayIn& r)
    {
        const float* q = q9 + 9 * i;
        make_ray(q[0], q[1], q[2], q[3], q[4], q[5], r);
        tx = q[6]; ty = q[7]; tz = q[8];
        return true;
    }
    bool decided(const RayIn& r, float best) const { return occlusion_decided(r, best, tx, ty, tz); }
    void done(size_t i, const RayIn& r, int id, float t) const
    {
        const float* q = q9 + 9 * i;
        bool vis = id < 0;
        if (!vis) {
            const float ex = (r.ox + r.dx * t) - q[6], ey = (r.oy + r.dy * t) - q[7], ez = (r.oz + r.dz * t) - q[8];
            vis = !(ex < -WRT_EPS) && !(ex > WRT_EPS) && !(ey < -WRT_EPS) && !(ey > WRT_EPS) && !(ez < -WRT_EPS) && !(ez > WRT_EPS);
        }
        occluded[i] = vis ? 0 : 1;
    }
};

struct WsScene { SceneLayout L; };

// One block of 128 lane-threads runs scheduler `sched` (2 = vote, 3 = pooled) over n work items.
template <bool PRUNED, class Src>
static void ws_run(const DevSceneView& sc, const Src& src0, size_t n, int sched)
{
    static PoolSmem pool_smem[4];
    std::vector<float4> scratch((size_t)4 * kPoolStack * kPoolRays);
    unsigned long long counter = 0;
    WsWarp warps[4];
    WsBarrier block_bar(128);
    ws_block_bar = &block_bar;
    std::vector<std::thread> th;
    for (unsigned tid = 0; tid < 128; tid++)
        th.emplace_back([&, tid] {
            threadIdx.x = tid; blockIdx.x = 0;
            ws_warp = &warps[tid >> 5]; ws_lane = tid & 31;
            Src src = src0;                                   // per-thread copy, like a kernel argument
            if (sched == 3) trace_pooled<PRUNED>(sc, src, &counter, n, pool_smem[tid >> 5], scratch.data() + (size_t)(tid >> 5) * kPoolStack * kPoolRays);
            else trace_persistent_vote<PRUNED>(sc, src, &counter, n);
        });
    for (auto& t : th) t.join();
}

extern "C" {

int ws_scene_create(const wrt_scene_desc* d, void** out, char* err256)
{
    WsScene* h = new WsScene();
    std::string err;
    if (!build_layout(d, h->L, err)) { if (err256) snprintf(err256, 256, "%s", err.c_str()); delete h; return 1; }
    layout_point_view_at_host(h->L);
    *out = h;
    return 0;
}

void ws_scene_destroy(void* h) { delete (WsScene*)h; }

void ws_trace_closest(void* hv, const wrt_ray* rays, size_t n, int pruned, int sched, int32_t* prim, float* t)
{
    const DevSceneView& sc = ((WsScene*)hv)->L.view;
    WsClosestSrc src = { rays, prim, t };
    if (pruned) ws_run<true>(sc, src, n, sched); else ws_run<false>(sc, src, n, sched);
}

// whole-warp traversals of last rays since the last call: finished that way / handed back to the ordinary rounds
void ws_par_stats(unsigned long long* out2) { out2[0] = g_par_ok.exchange(0); out2[1] = g_par_fallback.exchange(0); }

void ws_trace_occluded(void* hv, const float* q9, size_t n, int pruned, int sched, uint8_t* occ)
{
    const DevSceneView& sc = ((WsScene*)hv)->L.view;
    WsOccludedSrc src = { q9, occ, 0.f, 0.f, 0.f };
    if (pruned) ws_run<true>(sc, src, n, sched); else ws_run<false>(sc, src, n, sched);
}

}  // extern "C"
