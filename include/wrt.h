/*
 * wrt.h — C ABI of the B200-native render core for Winmad's raytracer.
 *
 * This is the drop-in boundary for ONE hot path of the reference renderer
 * (R = /root/reference/Winmad-s-raytracer-v1.0): KD-tree ray traversal with ray/triangle and
 * ray/sphere intersection, driven by the path-tracing and bidirectional-path-tracing bounce loops.
 * The reference has no FFI of its own; the seam is the C++ class boundary listed below, and each
 * entry point names the reference interface it replaces.  INTEGRATION.md shows the host shim a
 * maintainer of the reference would add (a SurfaceIntegrator subclass calling these functions).
 *
 * Conventions: plain pointers and sizes only; the library copies what it is given (callers keep
 * their arrays); all calls are synchronous unless they take a stream; return 0 (WRT_OK) or an
 * error code with text in wrt_last_error().  A miss is prim == -1 (the reference returns NULL).
 * There is no CPU fallback: without a usable CUDA device every compute call fails with
 * WRT_ERR_NO_DEVICE.
 */
#ifndef WRT_H
#define WRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WRT_OK 0
#define WRT_ERR_INVALID 1   /* bad argument / malformed scene description */
#define WRT_ERR_CUDA 2      /* CUDA runtime error (text in wrt_last_error) */
#define WRT_ERR_IO 3        /* scene / OBJ / image file problem */
#define WRT_ERR_NOMEM 4
#define WRT_ERR_NO_DEVICE 5 /* no CUDA device: the product path never falls back to the CPU */

#define WRT_PRIM_TRIANGLE 0
#define WRT_PRIM_SPHERE 1

#define WRT_TRAVERSE_EXACT 0  /* every leaf the ray pierces, every primitive in it (KDtreeAccel.cpp:321-385) */
#define WRT_TRAVERSE_PRUNED 1 /* same visit order; skips sub-trees that cannot change the result */

/* A ray exactly as the reference holds it AFTER the Ray constructor ran (dir already divided by
 * its length; tmin = 0, tmax = INF = 1e7).  R/src/geometry/ray.h:6-32. */
typedef struct { float ox, oy, oz, dx, dy, dz, tmin, tmax; } wrt_ray;

/* The reference's pointer KD-tree (KDtreeAccelNode, R/src/scene/KDtreeAccel.h:26-45) written out
 * as arrays.  Node 0 is the root; nodes may be in any order.  A node is a leaf iff axis == -1
 * (that is the test traverse makes, KDtreeAccel.cpp:325).  Leaf i owns refs[first_ref[i] ..
 * first_ref[i]+n_ref[i]) — primitive indices (position in Scene::objs) in objlist order. */
typedef struct {
    int32_t n_nodes;
    const int32_t* axis;      /* [n_nodes] 0,1,2 or -1 */
    const float* split;       /* [n_nodes] splitPlane (interior nodes) */
    const int32_t* left;      /* [n_nodes] child indices (interior nodes) */
    const int32_t* right;
    const int32_t* first_ref; /* [n_nodes] (leaves) */
    const int32_t* n_ref;     /* [n_nodes] (leaves) objNum */
    int64_t n_refs;
    const int32_t* refs;      /* [n_refs] */
    float root_box[6];        /* root->box: l.xyz, r.xyz (KDtreeAccel.cpp:48-53) */
} wrt_kdtree;

/* Scene::objs / materials / lights (R/src/scene/scene.h:21-31) as arrays. */
typedef struct {
    int32_t n_prims;
    const int32_t* prim_kind;  /* [n_prims] WRT_PRIM_* */
    const float* prim_data;    /* [n_prims*9] triangle: p0 p1 p2; sphere: centre xyz, radius, 5 unused */
    const int32_t* prim_matid; /* [n_prims] >0 material, <0 emitter -(light+1), 0 = "no material" */
    int32_t n_materials;
    const float* materials;    /* [n_materials*11] diffuse rgb, phong rgb, phongExp, specular rgb, index */
    int32_t n_lights;
    const float* lights;       /* [n_lights*12] AreaLight ctor args: p0 p1 p2, intensity rgb (light.h:90-104) */
    wrt_kdtree tree;
} wrt_scene_desc;

/* What the device needs of Camera (R/src/scene/camera.h:8-27) after Camera::setup ran on the host. */
typedef struct {
    float pos[3];
    float forward[3];
    float image_plane_dist;
    float x_res, y_res;         /* Camera::xResolution / yResolution */
    float raster_to_world[16];  /* rasterToWorld.m, row major */
    float world_to_raster[16];  /* worldToRaster.m */
} wrt_camera;

typedef struct {
    int32_t width, height;     /* film size: SurfaceIntegrator::width/height */
    int32_t spp;               /* samplesPerPixel: defines the stratification grid (sampler.cpp:28-42) */
    int32_t max_depth;         /* PathIntegrator::maxTracingDepth */
    uint32_t seed;             /* counter-based RNG key (the reference's MT19937 stream is not reproduced) */
    int32_t sample_first;      /* this call renders samples k = sample_first + i*sample_stride < spp */
    int32_t sample_stride;     /* 0 is treated as 1 */
    float film_scale;          /* multiplies every sample; 0 means 1/spp (surfaceIntegrator.cpp:45) */
} wrt_pt_params;

typedef struct {
    int32_t width, height;     /* must be square, as in the reference (bidirPathTracing.cpp:29-45) */
    int32_t iterations;        /* BidirPathTracing::iterations */
    int32_t min_path_length, max_path_length, control_length; /* reference: 0, 10, 3 */
    uint32_t seed;
    int32_t iter_first, iter_stride; /* this call runs iterations it = iter_first + i*iter_stride < iterations */
    float film_scale;          /* 0 means 1/iterations (bidirPathTracing.cpp:45) */
    int32_t transpose_output;  /* 1: return the film as outputImage shows it (transposed, :31-44) */
} wrt_bdpt_params;

typedef struct {
    uint64_t closest_rays;     /* Scene::intersect(ray,inter) equivalents traced since creation/reset */
    uint64_t shadow_rays;      /* Scene::occluded equivalents */
    uint64_t samples;          /* camera paths started */
    uint64_t kernel_launches;  /* kernels of this library launched */
    /* filled only by wrt_trace_count_visits (reference-semantics work per batch): */
    uint64_t inner_visits, leaf_visits, tri_tests, sphere_tests;
    double last_render_ms;     /* device time of the last wrt_render_* call (CUDA events) */
    double last_trace_ms;      /* device time of the last wrt_trace_* kernel */
    /* per-stage device time of the last wrt_render_* call (CUDA events around every launch): */
    double extend_ms, shade_ms, shadow_ms;
    uint64_t extend_launches;  /* closest-hit (extend) kernel launches in the last render */
    uint64_t extend_rays;      /* rays those launches traced */
    /* multi-GPU renders (wrt_init with more than one device): */
    double reduce_ms;          /* device time of the film exchange (peer films summed onto device 0) of the last render */
    uint64_t devices_used;     /* devices the last render ran on */
} wrt_stats;

typedef struct wrt_scene wrt_scene;           /* device-resident scene */
typedef struct wrt_host_scene wrt_host_scene; /* host-side scene: what Scene::init leaves behind */

/* ---- library / device ------------------------------------------------------------------------- */
const char* wrt_version(void);
const char* wrt_last_error(void);
int wrt_device_count(int* count);
int wrt_set_device(int ordinal);               /* one process per GPU: call once with LOCAL_RANK */
/* Multi-GPU INSIDE one process (the reference is one process: `ToT <scene> <image> -p`, R/src/main.cpp:29-97).
 * wrt_init names the devices (device_ids == NULL: 0 .. n_gpus-1; n_gpus <= 0: every visible device) and enables peer access
 * between them.  Every wrt_scene created afterwards is replicated on all of them, and ONE wrt_render_pt / wrt_render_whitted /
 * wrt_render_bdpt call (host-film variants) drives them all: samples (iterations) are dealt round-robin to the devices with
 * the 1-device RNG keys, every device runs its wavefront on its own host thread, and the films — pre-scaled by 1/spp as
 * SurfaceIntegrator::render's film->scale (surfaceIntegrator.cpp:45) — are summed onto the first device by one kernel that
 * reads the peers over NVLink.  The image equals the 1-device image up to float summation order.  Level-1 queries run on
 * the first device.  Without wrt_init the library uses the current device only. */
int wrt_init(int n_gpus, const int* device_ids);
int wrt_shutdown(void);                        /* forget the device set (scenes are destroyed by their owners) */

/* ---- host side: the product's own restatement of Scene::init -------------------------------------
 * (not needed by a reference-side shim, which already has Scene, Camera and KDtreeAccel objects) */
/* Scene::loadScene(char*) — XML .scene + Wavefront OBJ (R/src/scene/scene.cpp:259-467). */
int wrt_host_scene_load(const char* scene_file, wrt_host_scene** out);
/* Programmatic equivalent of the addMaterial/addGeometry/addLight/camera.setup calls the loader makes. */
int wrt_host_scene_from_arrays(int32_t n_materials, const float* materials11,
                               int32_t n_prims, const int32_t* prim_kind, const float* prim_data9,
                               const int32_t* prim_matid, int32_t n_lights, const float* lights12,
                               const float* cam_pos3_fwd3_up3_xres_yres_fov, wrt_host_scene** out);
/* KDtreeAccel::init + buildTree (KDtreeAccel.cpp:12-307): reproduces the reference tree exactly. */
int wrt_host_scene_build_kdtree(wrt_host_scene* hs);
/* Borrow the arrays (valid until wrt_host_scene_free). */
int wrt_host_scene_desc(const wrt_host_scene* hs, wrt_scene_desc* out);
int wrt_host_scene_camera(const wrt_host_scene* hs, wrt_camera* out);
/* sceneSphere: centre xyz, radius, invSceneRadiusSqr (scene.cpp:483-487). */
int wrt_host_scene_sphere(const wrt_host_scene* hs, float out5[5]);
void wrt_host_scene_free(wrt_host_scene* hs);
/* Flattened-scene cache on disk (KD build of multi-million-primitive scenes takes minutes). */
int wrt_host_scene_save(const wrt_host_scene* hs, const char* path);
int wrt_host_scene_load_cache(const char* path, wrt_host_scene** out);
/* Camera::setup (R/src/scene/camera.cpp:3-29). */
int wrt_camera_setup(const float pos[3], const float forward[3], const float up[3],
                     float x_res, float y_res, float horizontal_fov_deg, wrt_camera* out);
/* Camera::generateRay (camera.cpp:37-42) on the host, for building ray batches. */
int wrt_camera_generate_rays(const wrt_camera* cam, const float* xy, size_t n, wrt_ray* rays);
/* Ray(origin, dir) constructor (ray.h:14-16): 6 floats in, wrt_ray out. */
int wrt_make_rays(const float* origin_dir6, size_t n, wrt_ray* rays);
/* ImageFilm::outputImage(filename, scale, gamma) (film.cpp:39-64): scale, clamp, gamma, 8-bit; writes
 * binary PPM (or PNG-less BMP by extension ".bmp").  film is H x W x 3 linear floats. */
int wrt_film_write(const char* path, const float* film_hw3, int32_t width, int32_t height,
                   float scale, float gamma);

/* ---- device scene -------------------------------------------------------------------------------- */
/* Uploads primitives, the flattened KD-tree (re-laid out as 32-byte nodes with sub-tree bounds),
 * materials and lights to the current device. */
int wrt_scene_create(const wrt_scene_desc* desc, wrt_scene** out);
void wrt_scene_destroy(wrt_scene* sc);
int wrt_scene_set_traversal(wrt_scene* sc, int mode);  /* WRT_TRAVERSE_*; default PRUNED */
/* Work accounting: on = 1: the integrators' trace kernels run the EXACT (reference-semantics) traversal and
 * add their per-ray visit counts to wrt_stats (inner_visits ... sphere_tests); on = 2: same counters for the
 * PRUNED traversal, i.e. the work the production kernels really do. */
int wrt_scene_set_counting(wrt_scene* sc, int on);
int wrt_get_stats(wrt_scene* sc, wrt_stats* out);
int wrt_reset_stats(wrt_scene* sc);

/* ---- level 1: ray queries (R/src/scene/scene.h:44-50) ------------------------------------------------ */
/* One call traces at most 2^31 - 1 rays / queries (WRT_ERR_INVALID beyond that: split the batch). */
/* Geometry* Scene::intersect(const Ray&, Intersection&)   scene.cpp:21-43.
 * prim[i] = index in Scene::objs or -1; t[i] = inter.t (INF=1e7 on a miss). Host buffers. */
int wrt_trace_closest(wrt_scene* sc, const wrt_ray* rays, size_t n, int32_t* prim, float* t);
/* Same, plus the rest of Intersection (p, n, inside, matId); any output may be NULL. */
int wrt_trace_closest_full(wrt_scene* sc, const wrt_ray* rays, size_t n, int32_t* prim, float* t,
                           float* p3, float* n3, int32_t* inside, int32_t* matid);
/* bool Scene::intersect(const Ray&)   scene.cpp:45-53. */
int wrt_trace_any(wrt_scene* sc, const wrt_ray* rays, size_t n, uint8_t* hit);
/* Real Scene::shadowRayTest(const Ray&, const Vector3& p)   scene.cpp:55-69: visible[i] = 1.0f / 0.0f. */
int wrt_trace_shadow(wrt_scene* sc, const wrt_ray* rays, const float* target3, size_t n, float* visible);
/* bool Scene::occluded(p1, dir, p2)   scene.cpp:71-81: 9 floats per query; dir is normalised by the
 * Ray constructor inside, exactly as the reference does. */
int wrt_trace_occluded(wrt_scene* sc, const float* p1_dir_p2, size_t n, uint8_t* occluded);
/* Device-pointer variants (inputs/outputs already resident in HBM; stream = cudaStream_t or NULL).  Asynchronous on
 * `stream`.  Calls on DIFFERENT streams of one scene may overlap: every stream gets its own work counter and traversal
 * scratch (at most 16 distinct streams per scene); calls on one stream are stream-ordered as usual. */
int wrt_trace_closest_dev(wrt_scene* sc, const wrt_ray* d_rays, size_t n, int32_t* d_prim, float* d_t,
                          void* stream);
int wrt_trace_occluded_dev(wrt_scene* sc, const float* d_p1_dir_p2, size_t n, uint8_t* d_occluded,
                           void* stream);
/* Reference-semantics work counters for a batch (full traversal): fills inner_visits, leaf_visits,
 * tri_tests, sphere_tests in wrt_stats. Host rays. */
int wrt_trace_count_visits(wrt_scene* sc, const wrt_ray* rays, size_t n);

/* ---- level 2: integrators (R/src/surfaceIntegrator/surfaceIntegrator.h:14-34) ---------------------- */
/* PathIntegrator: SurfaceIntegrator::render + PathIntegrator::raytracing
 * (surfaceIntegrator.cpp:14-46, pathIntegrator.cpp:29-148).  film: H x W x 3 floats, film[i][j] as
 * ImageFilm::color[i][j] after film->scale(1/spp). */
int wrt_render_pt(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* film_hw3);
/* Device film (accumulated into: caller zeroes it); no host copy.  The work is ordered after everything queued on
 * `stream` so far; the call itself returns when the render has finished (it is NOT asynchronous).
 * All wrt_render_* calls on one wrt_scene must come from one host thread at a time. */
int wrt_render_pt_dev(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* d_film_hw3,
                      void* stream);
/* WhittedIntegrator: SurfaceIntegrator::render + WhittedIntegrator::raytracing (surfaceIntegrator.cpp:14-46,
 * whitted.cpp:17-113).  Same parameters as path tracing: max_depth = MAX_TRACING_DEPTH,
 * one light sample per hit.  Pixels the reference leaves NaN (a light seen from its back side, unoccluded) are NaN. */
int wrt_render_whitted(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* film_hw3);
int wrt_render_whitted_dev(wrt_scene* sc, const wrt_camera* cam, const wrt_pt_params* p, float* d_film_hw3,
                           void* stream);
/* BidirPathTracing::render (bidirPathTracing.cpp:23-265) with the shipped controlLength gating. */
int wrt_render_bdpt(wrt_scene* sc, const wrt_camera* cam, const wrt_bdpt_params* p, float* film_hw3);
int wrt_render_bdpt_dev(wrt_scene* sc, const wrt_camera* cam, const wrt_bdpt_params* p,
                        float* d_film_hw3, void* stream);

/* ---- film resolve on the device ----------------------------------------------------- */
/* The per-pixel pipeline of ImageFilm::outputImage (film.cpp:44-61, color.h:47-75): scale, clamp to [0,1],
 * pow(1/gamma), (unsigned char)(c * 255.0).  d_film: H x W x 3 floats, d_rgb: H x W x 3 bytes (R,G,B). */
int wrt_film_resolve_dev(const float* d_film_hw3, int32_t width, int32_t height, float scale, float gamma,
                         uint8_t* d_rgb_hw3, void* stream);

/* ---- diagnostics: how parity with the reference is checked on the device ---------------------------------- */
/* The shading math of the bounce loops evaluated on the device, one reference function per call (`what`):
 *   0 BSDF::f   1 BSDF::sample   2 BSDF::pdf           (R/src/material/bsdf.cpp:102-335, fresnel.cpp:3-30)
 *   3 AreaLight::illuminance   4 ::emit   5 ::getRadiance   (R/src/scene/light.cpp:4-100)
 *   6 fresnelDielectric   7 sampleTriangle / sampleCosHemisphere / samplePowerCosHemisphere + pdfs (sampler.cpp)
 *   8 sample k of pixel (i, j): sampleRectangleStratified + Camera::generateRay (surfaceIntegrator.cpp:26-34; iparam = spp)
 * `in` / `out` hold n fixed-size records; the record layouts are documented in csrc/shading_kat.cuh and are the ones
 * oracle/ref_harness.cpp::ref_shading_batch uses for the reference's own functions.  Host buffers. */
int wrt_debug_shading(wrt_scene* sc, const wrt_camera* cam, int what, int iparam, const float* in, size_t n, float* out);
/* RNG replay: installs (copies) a tape of random floats; while installed, wrt_render_pt / wrt_render_bdpt draw the
 * numbers of sample (pixel p, sample k) from tape[(p*spp + k)*stride ...] (PT) and of light / camera path p of iteration
 * it from tape[((it*W*H + p)*2 + {0,1})*stride ...] (BDPT) instead of the counter-based generator.  With the tape the
 * reference recorded for the same samples (oracle/ref_harness.cpp) both renderers follow the same paths, and the films
 * can be compared per pixel without Monte-Carlo noise.  tape == NULL removes it. */
int wrt_debug_set_rng_tape(wrt_scene* sc, const float* tape, size_t n_floats, uint32_t stride);

#ifdef __cplusplus
}
#endif
#endif /* WRT_H */
