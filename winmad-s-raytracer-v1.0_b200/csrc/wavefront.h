// Persistent wavefront state shared by the PT and BDPT integrators: a pool of path slots in HBM,
// two index queues (current / next bounce), a shadow-ray queue and a device counter block.
#pragma once
#include "dev_scene.h"
#include "shading.cuh"

namespace wrt {

// counter block indices (unsigned long long each)
enum {
    WF_WORK = 0,        // dynamic work fetch: extend kernel
    WF_WORK2 = 1,       //                     shade kernel
    WF_WORK3 = 2,       //                     shadow kernel
    WF_NEXT_COUNT = 3,  // entries appended to the next-bounce queue
    WF_SHADOW_COUNT = 4,
    WF_AUX_COUNT = 5,   // BDPT direct-illumination entries
    WF_WORK4 = 6,      // BDPT DI kernel work fetch; PT: regenerated camera rays appended from the back of the queue
    WF_AUX2_COUNT = 7,  // BDPT: BSDF-sampled rays traced by the DI kernel
    WF_PER_ITER = 8,    // counters [0, WF_PER_ITER) are zeroed before every iteration
    WF_NEXT_SAMPLE = 8, // next camera sample to hand out (path regeneration)
    WF_LIGHT_VERTS = 9,
    WF_VISITS = 10,     // +0 inner, +1 leaf, +2 tri, +3 sphere (counting mode)
    WF_COUNTERS = 16
};

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
    float* film; size_t film_floats;    // library-owned device film for host-buffer entry points
    void* bdpt;                         // BDPT-only buffers (bdpt_wavefront.cu)
    size_t bdpt_bytes;
    void* whitted; size_t whitted_bytes; // Whitted-only pending lists (pt_wavefront.cu)
    cudaEvent_t* ev; int n_ev;          // stage-timing events (4 per iteration)
    void* trav_scratch; size_t trav_scratch_bytes;   // pooled scheduler's traversal stacks for this sub-pool's launches
    cudaStream_t stream;                // this sub-pool's own stream (PT runs sub-pools concurrently)
    cudaEvent_t join_ev;
};

namespace wrt {
int wavefront_get(wrt_scene* sc, int capacity, wrt_wavefront** out);
int wavefront_get_slot(wrt_scene* sc, int slot, int capacity, wrt_wavefront** out);
int wavefront_film(wrt_scene* sc, size_t floats, float** out);
int persistent_grid_for(const void* kernel, int block);
int wavefront_events(wrt_wavefront* wf, int n);
void wavefront_sum_stage_times(wrt_scene* sc, wrt_wavefront* wf, int iters_timed);
}
#endif
