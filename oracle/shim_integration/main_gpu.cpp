// TEST INFRASTRUCTURE.  The reference's command line (R/src/main.cpp:29-97) reduced to the integrators on this seam, with the
// three branches INTEGRATION.md adds: `ToT_gpu <scene> <image> -gp | -gr | -gbpt | -gcheck | -p | -r | -bpt [parameters.para]`.
// Everything except render() is the UNMODIFIED reference (its loader, its KD builder, its camera, its film output, compiled
// from /root/reference by oracle/Makefile); render() goes through the shim (gpuIntegrator.h = the code block of
// INTEGRATION.md, byte for byte) into libwrt_b200.so.  tests/test_gpu_render.py::test_reference_with_shim runs it on a B200.
#include <cstdio>
#include <cstring>
#include <ctime>
#include "parameters.h"
#include "gpuIntegrator.h"
#include <vector>
#include <unordered_map>

Parameters para;
PathIntegrator pathIntegrator;
WhittedIntegrator whitted;
BidirPathTracing bidirPathTracing;
GpuPathIntegrator gpuPath;
GpuWhittedIntegrator gpuWhitted;
GpuBidirPathTracing gpuBidir;

// -gcheck: level-1 parity inside the reference's own process.  For every pixel centre: the reference's Scene::intersect on the
// camera ray, then on a reflected ray from the hit, and Scene::occluded towards the first light's first vertex — against the
// same batches through wrt_trace_closest / wrt_trace_occluded on the scene the shim uploaded (the reference's own objects and
// its own KD-tree).  Bar: Geometry* identity, t bit-identical, occlusion flags identical.  Prints and returns the mismatch count.
static int level1_check(GpuPathIntegrator& in)
{
    Scene& sc = in.scene;
    std::unordered_map<const Geometry*, int> index;
    for (size_t i = 0; i < sc.objs.size(); i++) index[sc.objs[i]] = (int)i;
    std::vector<wrt_ray> rays;
    std::vector<int> want_prim; std::vector<float> want_t;
    std::vector<float> occ_q; std::vector<unsigned char> want_occ;
    auto push_ray = [&](const Ray& r) {
        wrt_ray w = { r.origin.x, r.origin.y, r.origin.z, r.dir.x, r.dir.y, r.dir.z, r.tmin, r.tmax };
        rays.push_back(w);
        Intersection inter;
        Ray copy = r;
        Geometry* g = sc.intersect(copy, inter);
        want_prim.push_back(g ? index.at(g) : -1);
        want_t.push_back(g ? inter.t : INF);
        return g ? inter : Intersection();
    };
    Vector3 lp(0, 0, 0);
    for (size_t i = 0; i < sc.objs.size(); i++) if (sc.objs[i]->getMatId() < 0) { lp = ((Triangle*)sc.objs[i])->p0; break; }
    for (int i = 0; i < in.height; i++)
        for (int j = 0; j < in.width; j++) {
            Ray r = sc.camera.generateRay((Real)j, (Real)i);
            const size_t at = want_prim.size();
            Intersection inter = push_ray(r);
            if (want_prim[at] < 0) continue;
            Vector3 d = getReflectDir(-r.dir, inter.n);
            push_ray(Ray(inter.p + d * EPS, d));
            Vector3 wi = lp - inter.p; Real dist = wi.length(); wi = wi / dist;
            const float q[9] = { inter.p.x, inter.p.y, inter.p.z, wi.x, wi.y, wi.z, lp.x, lp.y, lp.z };
            occ_q.insert(occ_q.end(), q, q + 9);
            want_occ.push_back(sc.occluded(inter.p, wi, lp) ? 1 : 0);
        }
    std::vector<int32_t> prim(rays.size()); std::vector<float> t(rays.size()); std::vector<uint8_t> occ(want_occ.size());
    if (wrt_trace_closest(in.gpu.dev, rays.data(), rays.size(), prim.data(), t.data()) != WRT_OK ||
        wrt_trace_occluded(in.gpu.dev, occ_q.data(), occ.size(), occ.data()) != WRT_OK) {
        fprintf(stderr, "wrt: %s\n", wrt_last_error());
        return -1;
    }
    long bad = 0;
    for (size_t k = 0; k < rays.size(); k++)
        if (prim[k] != want_prim[k] || memcmp(&t[k], &want_t[k], 4) != 0) bad++;
    for (size_t k = 0; k < occ.size(); k++) if (occ[k] != want_occ[k]) bad++;
    printf("level-1 check: %zu closest-hit queries, %zu occlusion queries, %ld mismatches\n", rays.size(), occ.size(), bad);
    return (int)(bad > 0);
}

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
    else if (!strcmp(argv[3], "-gcheck")) { gpuPath.init(argv[1], para); return level1_check(gpuPath) == 0 ? 0 : 1; }
    else { printf("error!\n"); return 1; }
    clock_t end = clock();
    if (FILE* fp = fopen("time.txt", "w")) { fprintf(fp, "time = %d\n", (int)(end - start)); fclose(fp); }
    return 0;
}
