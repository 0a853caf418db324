// wrt_tot — the reference's command line for the two integrators this repository replaces:
//
//     wrt_tot <scene file> <output image (.ppm / .bmp)> -p | -r | -bpt [parameters.para]
//
// mirrors `ToT <scene> <image> <mode>` (R/src/main.cpp:29-97) for mode -p (PathIntegrator) and -bpt
// (BidirPathTracing): loads `src/parameters.para` (or the given file) exactly like
// Parameters::load_parameters (R/src/parameters.cpp:9-34), runs Scene::init on the host (our loader +
// reference-exact KD builder), renders on the GPU through the C ABI and writes the image like
// ImageFilm::outputImage (film.cpp:39-64; BDPT transposes and divides by `iterations` first,
// bidirPathTracing.cpp:29-46).  Timing goes to time.txt like the reference's (main.cpp:93-95).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/wrt.h"

struct Parameters {  // R/src/parameters.h
    int MAX_TRACING_DEPTH = 7, SAMPLES_PER_PIXEL = 1, SAMPLES_OF_LIGHT = 8, SAMPLES_OF_HEMISPHERE = 4;
    int WIDTH = 512, HEIGHT = 512, PHONG_POWER_INDEX = 5, POINT_LIGHT_NUM = 400;
    bool load(const char* filename)
    {
        FILE* fp = fopen(filename, "r");
        if (!fp) return false;
        int* fields[8] = { &MAX_TRACING_DEPTH, &SAMPLES_PER_PIXEL, &SAMPLES_OF_LIGHT, &SAMPLES_OF_HEMISPHERE,
                           &WIDTH, &HEIGHT, &PHONG_POWER_INDEX, &POINT_LIGHT_NUM };
        char str[1024];
        int k = 0;
        while (k < 8 && fscanf(fp, "%1023s", str) != EOF) {
            if (str[0] == '#' || str[0] == '\0') continue;   // read_int(): '#' tokens are comments
            *fields[k++] = atoi(str);
        }
        fclose(fp);
        return true;
    }
};

static int fail(const char* what)
{
    fprintf(stderr, "wrt_tot: %s: %s\n", what, wrt_last_error());
    return 1;
}

int main(int argc, char* argv[])
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s <scene> <image.ppm|.bmp> -p|-r|-bpt [parameters.para]\n", argv[0]);
        return 2;
    }
    Parameters para;
    const char* para_file = argc > 4 ? argv[4] : "src/parameters.para";
    if (!para.load(para_file)) fprintf(stderr, "wrt_tot: %s not found, using the shipped defaults\n", para_file);
    const auto t0 = std::chrono::steady_clock::now();

    wrt_host_scene* hs = nullptr;
    if (wrt_host_scene_load(argv[1], &hs)) return fail("scene load");
    if (wrt_host_scene_build_kdtree(hs)) return fail("kd build");
    wrt_scene_desc desc;
    wrt_camera cam;
    if (wrt_host_scene_desc(hs, &desc) || wrt_host_scene_camera(hs, &cam)) return fail("scene description");
    wrt_scene* sc = nullptr;
    if (wrt_scene_create(&desc, &sc)) return fail("scene upload");

    std::vector<float> film((size_t)para.WIDTH * para.HEIGHT * 3);
    int rc = 0;
    if (!strcmp(argv[3], "-p")) {
        wrt_pt_params p = { para.WIDTH, para.HEIGHT, para.SAMPLES_PER_PIXEL, para.MAX_TRACING_DEPTH, 0u, 0, 1, 0.f };
        rc = wrt_render_pt(sc, &cam, &p, film.data());
    } else if (!strcmp(argv[3], "-r")) {          // main.cpp:34-38 WhittedIntegrator
        wrt_pt_params p = { para.WIDTH, para.HEIGHT, para.SAMPLES_PER_PIXEL, para.MAX_TRACING_DEPTH, 0u, 0, 1, 0.f };
        rc = wrt_render_whitted(sc, &cam, &p, film.data());
    } else if (!strcmp(argv[3], "-bpt")) {
        // BidirPathTracing::init: min/max path length 0/10, 1 iteration, controlLength 3; outputImage transposes
        wrt_bdpt_params p = { para.WIDTH, para.HEIGHT, 1, 0, 10, 3, 0u, 0, 1, 0.f, 1 };
        rc = wrt_render_bdpt(sc, &cam, &p, film.data());
    } else {
        printf("error!\n");   // main.cpp:90-92
        rc = -1;
    }
    if (rc > 0) return fail("render");
    if (rc == 0 && wrt_film_write(argv[2], film.data(), para.WIDTH, para.HEIGHT, 1.f, 2.2f)) return fail("image write");

    wrt_stats st;
    wrt_get_stats(sc, &st);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (FILE* fp = fopen("time.txt", "w")) {
        fprintf(fp, "time = %.6f s (render %.3f ms on the GPU, %llu closest + %llu shadow rays, %llu samples)\n", secs,
                st.last_render_ms, (unsigned long long)st.closest_rays, (unsigned long long)st.shadow_rays,
                (unsigned long long)st.samples);
        fclose(fp);
    }
    wrt_scene_destroy(sc);
    wrt_host_scene_free(hs);
    return rc == 0 ? 0 : 1;
}
