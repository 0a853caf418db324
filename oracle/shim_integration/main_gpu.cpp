// TEST INFRASTRUCTURE.  The reference's command line (R/src/main.cpp:29-97) reduced to the integrators on this seam, with the
// three branches INTEGRATION.md adds: `ToT_gpu <scene> <image> -gp | -gr | -gbpt | -p | -r | -bpt [parameters.para]`.
// Everything except render() is the UNMODIFIED reference (its loader, its KD builder, its camera, its film output, compiled
// from /root/reference by oracle/Makefile); render() goes through the shim (gpuIntegrator.h = the code block of
// INTEGRATION.md, byte for byte) into libwrt_b200.so.  tests/test_gpu_render.py::test_reference_with_shim runs it on a B200.
#include <cstdio>
#include <cstring>
#include <ctime>
#include "parameters.h"
#include "gpuIntegrator.h"

Parameters para;
PathIntegrator pathIntegrator;
WhittedIntegrator whitted;
BidirPathTracing bidirPathTracing;
GpuPathIntegrator gpuPath;
GpuWhittedIntegrator gpuWhitted;
GpuBidirPathTracing gpuBidir;

int main(int argc, char* argv[])
{
    if (argc < 4) { fprintf(stderr, "usage: %s <scene> <image> -gp|-gr|-gbpt|-p|-r|-bpt [parameters.para]\n", argv[0]); return 2; }
    char default_para[] = "src/parameters.para";
    para.load_parameters(argc > 4 ? argv[4] : default_para);
    clock_t start = clock();
    if (!strcmp(argv[3], "-r")) { whitted.init(argv[1], para); whitted.render(); whitted.outputImage(argv[2]); }
    else if (!strcmp(argv[3], "-p")) { pathIntegrator.init(argv[1], para); pathIntegrator.render(); pathIntegrator.outputImage(argv[2]); }
    else if (!strcmp(argv[3], "-bpt")) { bidirPathTracing.init(argv[1], para); bidirPathTracing.render(); bidirPathTracing.outputImage(argv[2]); }
    else if (!strcmp(argv[3], "-gp"))   { gpuPath.init(argv[1], para);  gpuPath.render();  gpuPath.outputImage(argv[2]); }
    else if (!strcmp(argv[3], "-gr"))   { gpuWhitted.init(argv[1], para); gpuWhitted.render(); gpuWhitted.outputImage(argv[2]); }
    else if (!strcmp(argv[3], "-gbpt")) { gpuBidir.init(argv[1], para); gpuBidir.render(); gpuBidir.outputImage(argv[2]); }
    else { printf("error!\n"); return 1; }
    clock_t end = clock();
    if (FILE* fp = fopen("time.txt", "w")) { fprintf(fp, "time = %d\n", (int)(end - start)); fclose(fp); }
    return 0;
}
