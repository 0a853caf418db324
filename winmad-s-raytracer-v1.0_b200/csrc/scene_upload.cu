// wrt_scene_create / destroy and device management: uploads the layout built by scene_layout.cpp
// (32-byte KD nodes, 48-byte leaf records, primitives, materials, lights) to HBM.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "scene_layout.h"
#include "wavefront.h"

namespace wrt {

int cuda_fail(cudaError_t e, const char* what)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    set_error(buf);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? WRT_ERR_NO_DEVICE : WRT_ERR_CUDA;
}

int ensure_scratch(wrt_scene* sc, size_t in_bytes, size_t out_bytes)
{
    if (in_bytes > sc->scratch_in_bytes) {
        if (sc->d_scratch_in) cudaFree(sc->d_scratch_in);
        sc->d_scratch_in = nullptr; sc->scratch_in_bytes = 0;
        WRT_CUDA(cudaMalloc(&sc->d_scratch_in, in_bytes));
        sc->scratch_in_bytes = in_bytes;
    }
    if (out_bytes > sc->scratch_out_bytes) {
        if (sc->d_scratch_out) cudaFree(sc->d_scratch_out);
        sc->d_scratch_out = nullptr; sc->scratch_out_bytes = 0;
        WRT_CUDA(cudaMalloc(&sc->d_scratch_out, out_bytes));
        sc->scratch_out_bytes = out_bytes;
    }
    return WRT_OK;
}

}  // namespace wrt

using namespace wrt;

extern "C" {

int wrt_device_count(int* count)
{
    if (!count) { set_error("null argument"); return WRT_ERR_INVALID; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount"); }
    *count = n;
    return WRT_OK;
}

int wrt_set_device(int ordinal)
{
    WRT_CUDA(cudaSetDevice(ordinal));
    return WRT_OK;
}

// Uploads a built layout to the CURRENT device.
static int scene_create_on_current_device(SceneLayout& L, wrt_scene** out)
{
    std::vector<float4>& nodes = L.nodes; std::vector<float4>& recs = L.recs; std::vector<float4>& prims = L.prims;
    std::vector<DevMaterial>& mats = L.materials; std::vector<DevLight>& lights = L.lights;
    wrt_scene* sc = new wrt_scene();
    memset(sc, 0, sizeof *sc);
    sc->traversal_mode = WRT_TRAVERSE_PRUNED;
    sc->n_leaf_recs = L.n_recs;
    cudaGetDevice(&sc->device);
#define UP(dst, vec, T_) do { \
        size_t bytes__ = std::max<size_t>((vec).size(), 1) * sizeof(T_); \
        cudaError_t e__ = cudaMalloc(&(dst), bytes__); \
        if (e__ == cudaSuccess && !(vec).empty()) e__ = cudaMemcpy((dst), (vec).data(), (vec).size() * sizeof(T_), cudaMemcpyHostToDevice); \
        if (e__ != cudaSuccess) { int rc__ = cuda_fail(e__, "scene upload"); wrt_scene_destroy(sc); return rc__; } } while (0)
    {   // The root is node 0 and sibling pairs start at odd indices (breadth-first layout).  Placing node 0 at
        // byte 32 of the allocation makes every child pair one aligned 64-byte read.
        const size_t bytes = (nodes.size() + 2) * sizeof(float4);
        cudaError_t e = cudaMalloc(&sc->d_nodes, bytes);
        if (e == cudaSuccess) e = cudaMemset(sc->d_nodes, 0, 2 * sizeof(float4));
        if (e == cudaSuccess) e = cudaMemcpy((float4*)sc->d_nodes + 2, nodes.data(), nodes.size() * sizeof(float4), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { int rc = cuda_fail(e, "scene upload"); wrt_scene_destroy(sc); return rc; }
    }
    UP(sc->d_leaf_recs, recs, float4);
    UP(sc->d_prims, prims, float4);
    UP(sc->d_materials, mats, DevMaterial);
    UP(sc->d_lights, lights, DevLight);
#undef UP
    {
        cudaError_t e = cudaMalloc((void**)&sc->d_counters, 64 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(sc->d_counters, 0, 64 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&sc->ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&sc->ev1);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sc->ev_fork, cudaEventDisableTiming);
        if (e != cudaSuccess) { int rc = cuda_fail(e, "scene resources"); wrt_scene_destroy(sc); return rc; }
    }
    sc->view = L.view;
    DevSceneView& v = sc->view;
    v.nodes = (const float4*)sc->d_nodes + 2; v.leaf_recs = (const float4*)sc->d_leaf_recs;
    v.prims = (const float4*)sc->d_prims;
    v.materials = (const DevMaterial*)sc->d_materials; v.lights = (const DevLight*)sc->d_lights;
    *out = sc;
    return WRT_OK;
}

int wrt_scene_create(const wrt_scene_desc* d, wrt_scene** out)
{
    if (!d || !out) { set_error("wrt_scene_create: null argument"); return WRT_ERR_INVALID; }
    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            set_error("wrt_scene_create: no CUDA device (this library has no CPU path)");
            return WRT_ERR_NO_DEVICE;
        }
    }
    SceneLayout L;
    std::string err;
    if (!build_layout(d, L, err)) { set_error("wrt_scene_create: " + err); return WRT_ERR_INVALID; }
    const int n_multi = multi_device_count();
    if (n_multi <= 1) {
        if (n_multi == 1) WRT_CUDA(cudaSetDevice(multi_device(0)));
        return scene_create_on_current_device(L, out);
    }
    // wrt_init named several devices: the primary scene on the first, one replica on each of the others
    WRT_CUDA(cudaSetDevice(multi_device(0)));
    wrt_scene* primary = nullptr;
    int rc = scene_create_on_current_device(L, &primary);
    if (rc) return rc;
    for (int i = 1; i < n_multi; i++) {
        cudaError_t e = cudaSetDevice(multi_device(i));
        wrt_scene* r = nullptr;
        rc = e == cudaSuccess ? scene_create_on_current_device(L, &r) : cuda_fail(e, "cudaSetDevice");
        if (rc) { cudaSetDevice(multi_device(0)); wrt_scene_destroy(primary); return rc; }
        r->peer_of_primary = multi_peer_of_primary(i);
        primary->replica[primary->n_replicas++] = r;
    }
    WRT_CUDA(cudaSetDevice(multi_device(0)));
    *out = primary;
    return WRT_OK;
}

void wrt_scene_destroy(wrt_scene* sc)
{
    if (!sc) return;
    for (int i = 0; i < sc->n_replicas; i++) wrt_scene_destroy(sc->replica[i]);
    sc->n_replicas = 0;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(sc->device);
    wavefront_destroy(sc);
    cudaFree(sc->d_nodes); cudaFree(sc->d_leaf_recs); cudaFree(sc->d_prims);
    cudaFree(sc->d_materials); cudaFree(sc->d_lights);
    cudaFree(sc->d_scratch_in); cudaFree(sc->d_scratch_out); cudaFree(sc->d_counters); cudaFree(sc->d_trav_scratch);
    cudaFree(sc->d_rng_tape); cudaFree(sc->d_film);
    for (int i = 0; i < sc->n_trace_ctx; i++) { cudaFree(sc->trace_ctx[i].counter); cudaFree(sc->trace_ctx[i].scratch); }
    if (sc->stream) cudaStreamDestroy(sc->stream);
    if (sc->ev0) cudaEventDestroy(sc->ev0);
    if (sc->ev1) cudaEventDestroy(sc->ev1);
    if (sc->ev_fork) cudaEventDestroy(sc->ev_fork);
    delete sc;
    cudaSetDevice(prev_dev);
}

int wrt_scene_set_traversal(wrt_scene* sc, int mode)
{
    if (!sc || (mode != WRT_TRAVERSE_EXACT && mode != WRT_TRAVERSE_PRUNED)) { set_error("bad traversal mode"); return WRT_ERR_INVALID; }
    sc->traversal_mode = mode;
    for (int i = 0; i < sc->n_replicas; i++) sc->replica[i]->traversal_mode = mode;
    return WRT_OK;
}

int wrt_scene_set_counting(wrt_scene* sc, int on)
{
    if (!sc) { set_error("null scene"); return WRT_ERR_INVALID; }
    sc->counting = on == 2 ? 2 : (on ? 1 : 0);   // 1: EXACT (reference-semantics work), 2: PRUNED (this kernel's own work)
    for (int i = 0; i < sc->n_replicas; i++) sc->replica[i]->counting = sc->counting;
    return WRT_OK;
}

int wrt_get_stats(wrt_scene* sc, wrt_stats* out)
{
    if (!sc || !out) { set_error("null argument"); return WRT_ERR_INVALID; }
    *out = sc->stats;
    return WRT_OK;
}

int wrt_reset_stats(wrt_scene* sc)
{
    if (!sc) { set_error("null argument"); return WRT_ERR_INVALID; }
    memset(&sc->stats, 0, sizeof sc->stats);
    return WRT_OK;
}

}  // extern "C"
