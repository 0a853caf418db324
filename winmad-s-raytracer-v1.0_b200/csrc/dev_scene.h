// Device-resident scene layout (HBM) shared by all kernels.  See DESIGN.md "Data layout in HBM".
#pragma once
#include <cstdint>
#include <string>
#include "hd_compat.h"
#include "../../include/wrt.h"

namespace wrt {

#define WRT_EPS 1e-3f
#define WRT_INF 1e7f
#define WRT_STACK_DEPTH 32      /* ONE bound for every scheduler; deeper trees are refused by build_layout (reference: depMax + 5 <= 29 at N = 1e8) */
#define WRT_LEAF_TAG 3u
#define WRT_REC_SKIP 2          /* leaf record kind: conservative box of the next n records (kinds 0 / 1: triangle / sphere) */

// KD node, 32 bytes = one L2 sector, two 16-byte loads:
//   a.x  split plane (interior)            | first leaf record (leaf), as int bits
//   a.y  (child_pair_index << 2) | axis    | (n_records << 2) | 3
//   a.z, a.w, b.x : lo.x lo.y lo.z   conservative bounds of every primitive referenced below
//   b.y, b.z, b.w : hi.x hi.y hi.z
// Children of an interior node are adjacent: left = pair, right = pair + 1.  Nodes are laid out
// breadth-first, so the top levels of the tree are the first nodes of the array.
struct DevNodeHalf { float x, y, z, w; };

// Leaf record, 48 bytes, one per (leaf, primitive) reference, stored in leaf order so a leaf's
// primitives are contiguous:
//   triangle: r0 = p0.xyz, prim id   r1 = p0-p1 (A,B,C), 0      r2 = p0-p2 (D,E,F), kind=0
//   sphere:   r0 = c.xyz,  prim id   r1 = radius, box.l.xyz     r2 = box.r.xyz,     kind=1
//   skip:     r0 = lo.xyz, n         r1 = hi.xyz, 0             r2 = 0, 0, 0,       kind=2
// A skip record holds the conservative box of the n records that follow it (primitive records of consecutive list
// entries, possibly with nested skip records): PRUNED traversal jumps over them when the box is prunable.
// (p0-p1 and p0-p2 are the same float subtractions Triangle::hit performs first, triangle.cpp:24-30.)

struct DevMaterial {  // Material, R/src/material/material.h:7-31
    float diffuse[3];
    float phong[3];
    float phong_exp;
    float specular[3];
    float index;
    float pad;
};

struct DevLight {  // AreaLight, R/src/scene/light.h:82-129
    float p0[3], d1[3], d2[3];
    float fx[3], fy[3], fz[3];  // localFrame (z = normal)
    float intensity[3];
    float inv_area;
};

struct DevSceneView {  // passed to kernels by value
    const float4* __restrict__ nodes;      // 2 per node
    const float4* __restrict__ leaf_recs;  // 3 per leaf reference
    const float4* __restrict__ prims;      // 3 per primitive: (p0,matid) (p1,kind) (p2,0) | (c,matid) (r,kind..) ..
    const DevMaterial* __restrict__ materials;
    const DevLight* __restrict__ lights;
    int n_nodes, n_prims, n_materials, n_lights;
    int small_tree;          // < 512 nodes: rays are a handful of steps long, the plain per-thread loop wins
    float root_lo[3], root_hi[3];
    float sphere_center[3], sphere_radius, inv_sphere_radius_sqr;
};

}  // namespace wrt

#ifdef __CUDACC__
#define WRT_MAX_TRACE_STREAMS 16
#define WRT_MAX_DEVICES 16
struct wrt_trace_ctx {               // per caller stream: work counter + traversal-stack scratch (trace_kernels.cu)
    cudaStream_t stream;
    unsigned long long* counter;
    void* scratch; size_t scratch_bytes;
};
struct wrt_scene {
    wrt::DevSceneView view;
    int device;
    int traversal_mode;
    int counting;
    void* d_nodes; void* d_leaf_recs; void* d_prims; void* d_materials; void* d_lights;
    int64_t n_leaf_recs;
    wrt_stats stats;
    // scratch for host-buffer entry points (grown on demand)
    void* d_scratch_in; size_t scratch_in_bytes;
    void* d_scratch_out; size_t scratch_out_bytes;
    unsigned long long* d_counters;  // small device counter block
    void* d_trav_scratch; size_t trav_scratch_bytes;   // pooled scheduler's traversal stacks (API calls on `stream`)
    wrt_trace_ctx trace_ctx[WRT_MAX_TRACE_STREAMS]; int n_trace_ctx;   // device-pointer entry points: one per caller stream
    cudaStream_t stream;
    cudaEvent_t ev0, ev1, ev_fork;
    float* d_rng_tape; size_t rng_tape_floats; uint32_t rng_tape_stride;   // diagnostics: RNG replay (wrt_debug_set_rng_tape)
    // multi-GPU (wrt_init): the primary scene (device 0 of the set) owns one replica per further device
    struct wrt_scene* replica[WRT_MAX_DEVICES - 1]; int n_replicas;
    bool peer_of_primary;            // (replicas) device 0 of the set can read this device's memory directly (NVLink / PCIe P2P)
    float* d_film; size_t film_floats;   // library-owned device film of the host-buffer render calls
    struct wrt_wavefront* wf;        // lazily created integrator state (sub-pool 0; owns the film and BDPT buffers)
    struct wrt_wavefront* wf_extra[7]; // further PT sub-pools (each with its own stream), see pt_wavefront.cu
};

#include <nvtx3/nvToolsExt.h>
namespace wrt {
// NVTX range around the host side of a library call (tracing); free when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what);
#define WRT_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return wrt::cuda_fail(e__, #call); } while (0)
int ensure_scratch(wrt_scene* sc, size_t in_bytes, size_t out_bytes);
int ensure_trav_scratch(void** ptr, size_t* bytes);
void wavefront_destroy(wrt_scene* sc);
}  // namespace wrt
#endif  // __CUDACC__
