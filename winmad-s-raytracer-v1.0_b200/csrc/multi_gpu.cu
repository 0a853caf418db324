// Multi-GPU inside the library: wrt_init(n, ids) names the devices; every wrt_scene created afterwards is
// replicated on all of them; ONE wrt_render_pt / wrt_render_whitted / wrt_render_bdpt call then drives them all:
//   * samples k = g, g + N, ... (PT) / iterations it = g, g + N, ... (BDPT) go to device g — RNG keys depend on (pixel, global
//     sample index) only, so the N-device image is the 1-device image summed in another order;
//   * one host thread per device runs that device's wavefront loop on its own replica and its own streams;
//   * the rank films (already scaled by 1/spp while they are accumulated: the scale is fused into film_add) are combined on
//     device 0 by k_film_sum_peers, which reads the other devices' films DIRECTLY over NVLink (peer-mapped pointers, one
//     pass, float4) — the exchange step of SURVEY 8(e) / K6 as one kernel, no staging copies and no host round trip;
//     when peer access is not available the films are staged through cudaMemcpyPeer.
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "wavefront.h"

namespace wrt {

struct MultiState { int n; int dev[WRT_MAX_DEVICES]; bool peer0[WRT_MAX_DEVICES]; };
static MultiState g_multi = { 0, { 0 }, { false } };

int multi_device_count() { return g_multi.n; }
int multi_device(int i) { return g_multi.dev[i]; }

struct PeerFilms { const float* p[WRT_MAX_DEVICES]; };

__global__ void k_film_sum_peers(float* __restrict__ dst, PeerFilms peers, int n_peers, size_t n)
{
    const size_t n4 = n / 4;
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = d4[i];
        for (int k = 0; k < n_peers; k++) {
            const float4 v = reinterpret_cast<const float4*>(peers.p[k])[i];      // a load over NVLink
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        d4[i] = acc;
    }
    for (size_t i = 4 * n4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float acc = dst[i];
        for (int k = 0; k < n_peers; k++) acc += peers.p[k][i];
        dst[i] = acc;
    }
}

// Runs fn(replica g, g, n_active, device film of replica g) on one host thread per device, then sums the films onto
// replica 0's film.  `floats` = film size; n_units = spp (PT) / iterations (BDPT): devices beyond n_units stay idle.
int multi_render(wrt_scene* sc, size_t floats, int n_units, const std::function<int(wrt_scene*, int, int, float*)>& fn, float** film0)
{
    NvtxRange nvtx_range("wrt: multi-GPU render (replicas + film exchange)");
    const int n_dev = 1 + sc->n_replicas;
    const int n_active = std::max(1, std::min(n_dev, n_units));
    std::vector<int> rc(n_active, WRT_OK);
    std::vector<std::string> err(n_active);
    std::vector<float*> films(n_active, nullptr);
    const auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int g) {
        wrt_scene* rs = g == 0 ? sc : sc->replica[g - 1];
        cudaError_t e = cudaSetDevice(rs->device);
        if (e != cudaSuccess) { rc[g] = cuda_fail(e, "cudaSetDevice"); err[g] = wrt_last_error(); return; }
        float* d_film = nullptr;
        int r = wavefront_film(rs, floats, &d_film);
        if (r == WRT_OK) { e = cudaMemsetAsync(d_film, 0, floats * sizeof(float), rs->stream); if (e != cudaSuccess) r = cuda_fail(e, "film memset"); }
        if (r == WRT_OK) r = fn(rs, g, n_active, d_film);
        if (r == WRT_OK) { e = cudaStreamSynchronize(rs->stream); if (e != cudaSuccess) r = cuda_fail(e, "replica stream"); }
        rc[g] = r; films[g] = d_film;
        if (r != WRT_OK) err[g] = wrt_last_error();
    };
    std::vector<std::thread> th;
    for (int g = 1; g < n_active; g++) th.emplace_back(work, g);
    work(0);
    for (auto& t : th) t.join();
    WRT_CUDA(cudaSetDevice(sc->device));
    for (int g = 0; g < n_active; g++) if (rc[g] != WRT_OK) { set_error("device " + std::to_string(g) + ": " + err[g]); return rc[g]; }
    // ---- the exchange step: sum the rank films onto device 0 ---------------------------------------------------------
    if (n_active > 1) {
        bool all_peer = true;
        for (int g = 1; g < n_active; g++) all_peer = all_peer && sc->replica[g - 1]->peer_of_primary;
        WRT_CUDA(cudaEventRecord(sc->ev0, sc->stream));
        if (all_peer) {
            PeerFilms pf; memset(&pf, 0, sizeof pf);
            for (int g = 1; g < n_active; g++) pf.p[g - 1] = films[g];
            const int block = 256;
            const int grid = (int)std::min<size_t>((floats / 4 + block - 1) / block + 1, (size_t)148 * 8);
            k_film_sum_peers<<<grid, block, 0, sc->stream>>>(films[0], pf, n_active - 1, floats);
            WRT_CUDA(cudaGetLastError());
            sc->stats.kernel_launches += 1;
        } else {
            float* stage = nullptr;
            WRT_CUDA(cudaMalloc((void**)&stage, floats * sizeof(float)));
            for (int g = 1; g < n_active; g++) {
                WRT_CUDA(cudaMemcpyPeerAsync(stage, sc->device, films[g], sc->replica[g - 1]->device, floats * sizeof(float), sc->stream));
                PeerFilms pf; memset(&pf, 0, sizeof pf); pf.p[0] = stage;
                k_film_sum_peers<<<148 * 4, 256, 0, sc->stream>>>(films[0], pf, 1, floats);
                WRT_CUDA(cudaGetLastError());
                sc->stats.kernel_launches += 1;
            }
            WRT_CUDA(cudaStreamSynchronize(sc->stream));
            cudaFree(stage);
        }
        WRT_CUDA(cudaEventRecord(sc->ev1, sc->stream));
        WRT_CUDA(cudaStreamSynchronize(sc->stream));
        float ms = 0.f; cudaEventElapsedTime(&ms, sc->ev0, sc->ev1);
        sc->stats.reduce_ms = ms;
    } else sc->stats.reduce_ms = 0.0;
    // statistics of the replicas flow into the primary's
    for (int g = 1; g < n_active; g++) {
        wrt_stats& a = sc->stats; wrt_stats& b = sc->replica[g - 1]->stats;
        a.closest_rays += b.closest_rays; a.shadow_rays += b.shadow_rays; a.samples += b.samples; a.kernel_launches += b.kernel_launches;
        a.inner_visits += b.inner_visits; a.leaf_visits += b.leaf_visits; a.tri_tests += b.tri_tests; a.sphere_tests += b.sphere_tests;
        a.extend_ms += b.extend_ms; a.shade_ms += b.shade_ms; a.shadow_ms += b.shadow_ms;
        a.extend_launches += b.extend_launches; a.extend_rays += b.extend_rays;
        memset(&b, 0, sizeof b);
    }
    sc->stats.last_render_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    sc->stats.devices_used = n_active;
    *film0 = films[0];
    return WRT_OK;
}

}  // namespace wrt

using namespace wrt;

extern "C" {

int wrt_init(int n_gpus, const int* device_ids)
{
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0) { set_error("wrt_init: no CUDA device (this library has no CPU path)"); return WRT_ERR_NO_DEVICE; }
    if (n_gpus <= 0) n_gpus = visible;                          // "all of them"
    if (n_gpus > WRT_MAX_DEVICES) { set_error("wrt_init: more devices than the library supports"); return WRT_ERR_INVALID; }
    MultiState m; memset(&m, 0, sizeof m);
    for (int i = 0; i < n_gpus; i++) {
        const int d = device_ids ? device_ids[i] : i;
        if (d < 0 || d >= visible) { set_error("wrt_init: device id out of range"); return WRT_ERR_INVALID; }
        for (int j = 0; j < i; j++) if (m.dev[j] == d) { set_error("wrt_init: duplicate device id"); return WRT_ERR_INVALID; }
        m.dev[i] = d;
    }
    m.n = n_gpus;
    // peer access from device 0 of the set to every other (the film exchange reads peers from there) and back
    for (int i = 0; i < n_gpus; i++) {
        WRT_CUDA(cudaSetDevice(m.dev[i]));
        WRT_CUDA(cudaFree(0));                                  // create the context
    }
    m.peer0[0] = true;
    for (int i = 1; i < n_gpus; i++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m.dev[0], m.dev[i]);
        if (can) {
            cudaSetDevice(m.dev[0]);
            e = cudaDeviceEnablePeerAccess(m.dev[i], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            can = e == cudaSuccess;
            if (!can) cudaGetLastError();
        }
        m.peer0[i] = can != 0;
    }
    WRT_CUDA(cudaSetDevice(m.dev[0]));
    g_multi = m;
    return WRT_OK;
}

int wrt_shutdown(void)
{
    memset(&g_multi, 0, sizeof g_multi);
    return WRT_OK;
}

}  // extern "C"

namespace wrt {
bool multi_peer_of_primary(int i) { return g_multi.peer0[i]; }
}
