// TEST INFRASTRUCTURE ONLY — interposer for the RNG-tape recorder of the BDPT parity tests.
//
// BidirPathTracing::runIteration (R/src/surfaceIntegrator/bidirPathTracing.cpp:53-265) is one monolithic loop over all
// light paths and then all camera paths, drawing every random number from one MT19937 stream.  To know where in that
// stream each path starts WITHOUT editing the reference, this library defines the two member functions the loop calls
// at the start of every path — generateLightSample (:267-311) and generateCameraSample (:418-452) — under their own
// mangled names.  libwrt_ref.so calls them through its PLT, so when this library is loaded first with RTLD_GLOBAL the
// calls land here; each hook records the `stride` floats the stream holds at that point and forwards to the real
// function (addresses handed in by oracle/refpy.py from the reference library's own symbol table).
#include "surfaceIntegrator/bidirPathTracing.h"
#include <cstring>

namespace {
typedef void (*LightFn)(BidirPathTracing*, BidirPathState&);
typedef Vector3 (*CameraFn)(BidirPathTracing*, int, BidirPathState&);
LightFn g_real_light = 0;
CameraFn g_real_camera = 0;
float* g_tape = 0;            // [2 * paths][stride]; light path n -> slot 2n, camera path n -> slot 2n + 1
int g_stride = 0;
long long g_capacity = 0;     // paths the tape has room for
long long g_light_calls = 0, g_camera_calls = 0;
long long* g_pos = 0;         // optional: absolute stream position at the start of every path, same slot order
long long g_blocks = 0; int g_last_mti = -1;

void record(const RNG& rng, long long slot)
{
    if (!g_tape || slot >= 2 * g_capacity) return;
    const int mti = rng.mti % RNG::N;
    if (g_last_mti >= 0 && mti < g_last_mti) g_blocks++;          // the stream regenerated its 624-word block since the last path
    g_last_mti = mti;
    if (g_pos) g_pos[slot] = g_blocks * RNG::N + mti;
    RNG copy = rng;
    float* dst = g_tape + slot * (long long)g_stride;
    for (int k = 0; k < g_stride; k++) dst[k] = copy.randFloat();
}
}  // namespace

void hook_generate_light_sample(BidirPathTracing* self, BidirPathState& st)
    asm("_ZN16BidirPathTracing19generateLightSampleER14BidirPathState");
void hook_generate_light_sample(BidirPathTracing* self, BidirPathState& st)
{
    record(self->rng, 2 * g_light_calls);
    g_light_calls++;
    g_real_light(self, st);
}

Vector3 hook_generate_camera_sample(BidirPathTracing* self, int pathIndex, BidirPathState& st)
    asm("_ZN16BidirPathTracing20generateCameraSampleEiR14BidirPathState");
Vector3 hook_generate_camera_sample(BidirPathTracing* self, int pathIndex, BidirPathState& st)
{
    record(self->rng, 2 * g_camera_calls + 1);
    g_camera_calls++;
    return g_real_camera(self, pathIndex, st);
}

extern "C" {
void ref_hooks_set_real(void* light, void* camera) { g_real_light = (LightFn)light; g_real_camera = (CameraFn)camera; }
void ref_hooks_start(float* tape, int stride, long long paths, long long* pos)
{
    g_tape = tape; g_stride = stride; g_capacity = paths; g_pos = pos;
    g_light_calls = g_camera_calls = 0; g_blocks = 0; g_last_mti = -1;
}
void ref_hooks_stop(long long* light_calls, long long* camera_calls)
{
    if (light_calls) *light_calls = g_light_calls;
    if (camera_calls) *camera_calls = g_camera_calls;
    g_tape = 0; g_pos = 0; g_stride = 0; g_capacity = 0;
}
}
