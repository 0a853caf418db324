// Diagnostics behind the C ABI (include/wrt.h, "diagnostics" block): the shading math evaluated on the device for
// known-answer tests against the reference's own functions, and the RNG tape that lets the integrators replay the
// reference's random numbers sample for sample.
#include <cstring>
#include "shading_kat.cuh"

namespace wrt {

void fill_camera(const wrt_camera* c, DevCamera& d);

__global__ void k_debug_shading(DevSceneView sc, DevCamera cam, int what, int iparam, const float* __restrict__ in, size_t n,
                                float* __restrict__ out)
{
    const int is = shading_kat_in_stride(what), os = shading_kat_out_stride(what);
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        float a[13], o[13];
        for (int k = 0; k < is; k++) a[k] = in[(size_t)is * e + k];
        shading_kat(sc, cam, what, iparam, a, o);
        for (int k = 0; k < os; k++) out[(size_t)os * e + k] = o[k];
    }
}

}  // namespace wrt

using namespace wrt;

extern "C" {

int wrt_debug_shading(wrt_scene* sc, const wrt_camera* cam, int what, int iparam, const float* in, size_t n, float* out)
{
    if (!sc) { set_error("wrt_debug_shading: null scene"); return WRT_ERR_INVALID; }
    const int is = shading_kat_in_stride(what), os = shading_kat_out_stride(what);
    if (!is || (n && (!in || !out)) || (what == 8 && !cam)) { set_error("wrt_debug_shading: bad argument"); return WRT_ERR_INVALID; }
    if (n == 0) return WRT_OK;
    WRT_CUDA(cudaSetDevice(sc->device));
    // light / material indices are taken from the input records: range-check them here, the kernels index directly
    for (size_t e = 0; e < n; e++) {
        const float* a = in + (size_t)is * e;
        if (what <= 2 && (int)a[6] >= sc->view.n_materials) { set_error("wrt_debug_shading: material id out of range"); return WRT_ERR_INVALID; }
        if (what >= 3 && what <= 5 && ((int)a[0] < 0 || (int)a[0] >= sc->view.n_lights)) { set_error("wrt_debug_shading: light id out of range"); return WRT_ERR_INVALID; }
    }
    int rc = ensure_scratch(sc, n * is * sizeof(float), n * os * sizeof(float));
    if (rc) return rc;
    DevCamera dc; memset(&dc, 0, sizeof dc);
    if (cam) fill_camera(cam, dc);
    cudaStream_t st = sc->stream;
    WRT_CUDA(cudaMemcpyAsync(sc->d_scratch_in, in, n * is * sizeof(float), cudaMemcpyHostToDevice, st));
    const int block = 128;
    const int grid = (int)((n + block - 1) / block < 148 * 8 ? (n + block - 1) / block : 148 * 8);
    k_debug_shading<<<grid, block, 0, st>>>(sc->view, dc, what, iparam, (const float*)sc->d_scratch_in, n, (float*)sc->d_scratch_out);
    WRT_CUDA(cudaGetLastError());
    sc->stats.kernel_launches += 1;
    WRT_CUDA(cudaMemcpyAsync(out, sc->d_scratch_out, n * os * sizeof(float), cudaMemcpyDeviceToHost, st));
    WRT_CUDA(cudaStreamSynchronize(st));
    return WRT_OK;
}

int wrt_debug_set_rng_tape(wrt_scene* sc, const float* tape, size_t n_floats, uint32_t stride)
{
    if (!sc) { set_error("wrt_debug_set_rng_tape: null scene"); return WRT_ERR_INVALID; }
    WRT_CUDA(cudaSetDevice(sc->device));
    if (sc->d_rng_tape) { cudaFree(sc->d_rng_tape); sc->d_rng_tape = nullptr; }
    sc->rng_tape_floats = 0; sc->rng_tape_stride = 0;
    if (!tape || n_floats == 0) return WRT_OK;               // tape removed: back to the counter-based RNG
    if (stride == 0 || n_floats >= ((size_t)1 << 32)) { set_error("wrt_debug_set_rng_tape: stride must be > 0 and the tape shorter than 2^32 floats"); return WRT_ERR_INVALID; }
    WRT_CUDA(cudaMalloc((void**)&sc->d_rng_tape, n_floats * sizeof(float)));
    WRT_CUDA(cudaMemcpy(sc->d_rng_tape, tape, n_floats * sizeof(float), cudaMemcpyHostToDevice));
    sc->rng_tape_floats = n_floats; sc->rng_tape_stride = stride;
    return WRT_OK;
}

}  // extern "C"
