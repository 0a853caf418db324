// Persistent wavefront state shared by the PT and BDPT integrators: a pool of path slots in HBM,
// two index queues (current / next bounce), a shadow-ray queue and a device counter block.
#pragma once
#include "dev_scene.h"
#include "shading.cuh"

namespace wrt {

// Device counter block (unsigned long long each).  The wavefront loops are DEVICE-DRIVEN: every iteration (extend ->
// shade -> shadow [-> direct illumination]) reads the length of its input queue from the counter bank the previous
// iteration wrote, so the host can enqueue iterations back to back without reading anything back.
// Two banks alternate: iteration i writes bank i & 1 and reads bank (i & 1) ^ 1.
enum {
    // ---- per-iteration bank (WF_BANK entries) ----
    WF_WORK = 0,        // dynamic work fetch: extend kernel
    WF_WORK2 = 1,       //                     shade kernel
    WF_WORK3 = 2,       //                     shadow kernel
    WF_NEXT_COUNT = 3,  // entries appended to the front of the next-bounce queue (continuing paths)
    WF_SHADOW_COUNT = 4,
    WF_AUX_COUNT = 5,   // BDPT direct-illumination entries
    WF_GEN_COUNT = 6,   // PT: regenerated camera rays appended from the back of the next-bounce queue (BDPT: stays 0)
    WF_AUX2_COUNT = 7,  // BDPT: BSDF-sampled rays traced by the DI kernel
    WF_WORK4 = 8,       // BDPT DI kernel work fetch
    WF_CV_COUNT = 9,    // BDPT: camera-vertex records written by the camera shade kernel for the connection kernel
    WF_PAIR_COUNT = 11, // BDPT: (camera vertex, light vertex) pairs listed for the connection kernel
    WF_BANK = 12,
    // ---- persistent part, after the two banks ----
    WF_PERSIST = 2 * WF_BANK,
    WF_NEXT_SAMPLE = WF_PERSIST + 0, // next camera sample to hand out (path regeneration)
    WF_TOTAL_CLOSEST = WF_PERSIST + 1, // rays traced by the extend kernels since the render started
    WF_TOTAL_SHADOW = WF_PERSIST + 2,  // queries traced by the shadow / DI kernels
    WF_TOTAL_ITERS = WF_PERSIST + 3,   // iterations that had work
    WF_VISITS = WF_PERSIST + 4,     // +0 inner, +1 leaf, +2 tri, +3 sphere (counting mode)
    WF_COUNTERS = WF_PERSIST + 8
};

// Length of the queue an iteration consumes: continuing paths (front of the queue) + regenerated camera rays (back).
__host__ __device__ inline size_t wf_queue_n(const unsigned long long* prev) { return (size_t)(prev[WF_NEXT_COUNT] + prev[WF_GEN_COUNT]); }

struct PathPool {            // SoA over `capacity` slots
    wrt_ray* ray;            // 32 B, float4-aligned
    float4* weight_pdf;      // pathWeight rgb, lastPdf
    uint4* meta;             // pixel, rng.key, rng.ctr, length | last_specular << 16
    int32_t* hit_prim;
    float* hit_t;
};

struct ShadowQueue {         // one pending connection per entry
    float4* a;               // p1.xyz, contrib.r
    float4* b;               // dir.xyz, contrib.g
    float4* c;               // p2.xyz, contrib.b
    uint32_t* pixel;
};

}  // namespace wrt

#ifdef __CUDACC__
struct wrt_wavefront {
    int capacity;
    wrt::PathPool pool;
    uint32_t* queue[2];
    wrt::ShadowQueue shadow;
    unsigned long long* counters;       // device, WF_COUNTERS entries
    unsigned long long* h_counters;     // pinned host mirror
    void* bdpt;                         // BDPT-only buffers (bdpt_wavefront.cu)
    size_t bdpt_bytes;
    void* whitted; size_t whitted_bytes; // Whitted-only pending lists (pt_wavefront.cu)
    cudaEvent_t* ev; int n_ev;          // stage-timing events (4 per iteration)
    void* trav_scratch; size_t trav_scratch_bytes;   // pooled scheduler's traversal stacks for this sub-pool's launches
    cudaStream_t stream;                // this sub-pool's own stream (PT runs sub-pools concurrently)
    cudaEvent_t join_ev;
    cudaEvent_t poll_ev[2];             // device-driven loops: marks the arrival of a batch's counter bank in h_counters
    // PT: the shadow kernel of iteration i runs on its own stream, next to extend / shade of iteration i + 1 (it only adds
    // to the film), so an iteration's critical path is extend -> shade.  Needs a second shadow queue (iteration parity), a
    // second traversal scratch and two events per parity.
    wrt::ShadowQueue shadow2;
    void* trav_scratch2; size_t trav_scratch2_bytes;
    cudaStream_t shadow_stream;
    cudaEvent_t shaded_ev[2], shadowed_ev[2];
};

#include <functional>
namespace wrt {
// multi_gpu.cu
int multi_device_count();
int multi_device(int i);
bool multi_peer_of_primary(int i);
int multi_render(wrt_scene* sc, size_t floats, int n_units, const std::function<int(wrt_scene*, int, int, float*)>& fn, float** film0);
int render_pt_device(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* d_film, cudaStream_t st, bool whitted);
int render_bdpt_device(wrt_scene* sc, const wrt_camera* cam, const wrt_bdpt_params* p, float* d_film, cudaStream_t st);
int wavefront_get(wrt_scene* sc, int capacity, wrt_wavefront** out);
int wavefront_get_slot(wrt_scene* sc, int slot, int capacity, wrt_wavefront** out);
int wavefront_film(wrt_scene* sc, size_t floats, float** out);
int persistent_grid_for(const void* kernel, int block);
int wavefront_events(wrt_wavefront* wf, int n);
void wavefront_sum_stage_times(wrt_scene* sc, wrt_wavefront* wf, int iters_timed);
constexpr int kEvPerIter = 5;   // [0] before extend [1] after extend [2] after shade [3] before shadow [4] after shadow
}
#endif
