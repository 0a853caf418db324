// R/src/surfaceIntegrator/gpuIntegrator.h
#include "pathIntegrator.h"
#include "bidirPathTracing.h"
#include "whitted.h"
#include <wrt.h>
#include <unordered_map>

struct GpuSceneBridge {                    // flattens Scene -> wrt_scene_desc, owns the device scene
    wrt_scene* dev = nullptr;
    wrt_camera cam;
    std::vector<int32_t> kind, matid, axis, left, right, first_ref, n_ref, refs;
    std::vector<float> data, split, materials, lights;

    int flatten(KDtreeAccelNode* tr, const std::unordered_map<const Geometry*, int>& index) {
        int me = (int)axis.size();
        axis.push_back(tr->axis); split.push_back(tr->splitPlane);
        left.push_back(-1); right.push_back(-1); first_ref.push_back(-1); n_ref.push_back(tr->objNum);
        if (tr->axis == -1) {                               // traverse()'s own leaf test, KDtreeAccel.cpp:325
            first_ref[me] = (int)refs.size();
            for (int i = 0; i < tr->objNum; i++) refs.push_back(index.at(tr->objlist[i]));   // objlist order matters
        } else {
            int l = flatten(tr->left, index), r = flatten(tr->right, index);
            left[me] = l; right[me] = r;
        }
        return me;
    }

    bool upload(Scene& sc) {
        std::unordered_map<const Geometry*, int> index;
        for (size_t i = 0; i < sc.objs.size(); i++) {
            index[sc.objs[i]] = (int)i;
            float d[9] = {0};
            if (Triangle* t = dynamic_cast<Triangle*>(sc.objs[i])) {
                kind.push_back(WRT_PRIM_TRIANGLE);
                d[0]=t->p0.x; d[1]=t->p0.y; d[2]=t->p0.z; d[3]=t->p1.x; d[4]=t->p1.y; d[5]=t->p1.z; d[6]=t->p2.x; d[7]=t->p2.y; d[8]=t->p2.z;
            } else if (Sphere* s = dynamic_cast<Sphere*>(sc.objs[i])) {
                kind.push_back(WRT_PRIM_SPHERE);
                d[0]=s->center.x; d[1]=s->center.y; d[2]=s->center.z; d[3]=s->radius;
            } else return false;
            data.insert(data.end(), d, d + 9);
            matid.push_back(sc.objs[i]->getMatId());
        }
        for (const Material& m : sc.materials) {
            float v[11] = { m.diffuse.r, m.diffuse.g, m.diffuse.b, m.phong.r, m.phong.g, m.phong.b, m.phongExp,
                            m.specular.r, m.specular.g, m.specular.b, m.index };
            materials.insert(materials.end(), v, v + 11);
        }
        // AreaLight ctor arguments: the emitter triangles (matId = -(f+1)) carry the same vertices, scene.cpp:424-428
        for (size_t li = 0; li < sc.lights.size(); li++) {
            AreaLight* a = dynamic_cast<AreaLight*>(sc.lights[li]);
            if (!a) return false;                            // only AreaLight is reachable from .scene files
            for (size_t i = 0; i < sc.objs.size(); i++)
                if (sc.objs[i]->getMatId() == -((int)li + 1)) {
                    lights.insert(lights.end(), &data[9 * i], &data[9 * i] + 9);
                    float c[3] = { a->intensity.r, a->intensity.g, a->intensity.b };
                    lights.insert(lights.end(), c, c + 3);
                    break;
                }
        }
        flatten(sc.kdtreeAccel.root, index);
        wrt_scene_desc d = {};
        d.n_prims = (int32_t)kind.size(); d.prim_kind = kind.data(); d.prim_data = data.data(); d.prim_matid = matid.data();
        d.n_materials = (int32_t)sc.materials.size(); d.materials = materials.data();
        d.n_lights = (int32_t)sc.lights.size(); d.lights = lights.data();
        d.tree.n_nodes = (int32_t)axis.size(); d.tree.axis = axis.data(); d.tree.split = split.data();
        d.tree.left = left.data(); d.tree.right = right.data(); d.tree.first_ref = first_ref.data(); d.tree.n_ref = n_ref.data();
        d.tree.n_refs = (int64_t)refs.size(); d.tree.refs = refs.data();
        const AABB& rb = sc.kdtreeAccel.root->box;
        float box[6] = { rb.l.x, rb.l.y, rb.l.z, rb.r.x, rb.r.y, rb.r.z };
        memcpy(d.tree.root_box, box, sizeof box);
        Camera& c = sc.camera;
        float pos[3] = { c.pos.x, c.pos.y, c.pos.z }, fwd[3] = { c.forward.x, c.forward.y, c.forward.z };
        memcpy(cam.pos, pos, sizeof pos); memcpy(cam.forward, fwd, sizeof fwd);
        cam.image_plane_dist = c.imagePlaneDist; cam.x_res = c.xResolution; cam.y_res = c.yResolution;
        memcpy(cam.raster_to_world, c.rasterToWorld.m.m, 16 * sizeof(float));   // Camera::setup already ran on the host
        memcpy(cam.world_to_raster, c.worldToRaster.m.m, 16 * sizeof(float));
        // all GPUs of the box behind the one render() call: WRT_GPUS=8 ./ToT scene out -gp  (scene replicated, samples sharded,
        // films summed on device 0 over NVLink inside wrt_render_*; without it the library uses the current device)
        if (const char* g = getenv("WRT_GPUS")) if (wrt_init(atoi(g), nullptr) != WRT_OK) { fprintf(stderr, "wrt: %s\n", wrt_last_error()); return false; }
        if (wrt_scene_create(&d, &dev) != WRT_OK) { fprintf(stderr, "wrt: %s\n", wrt_last_error()); return false; }
        return true;
    }
};

class GpuPathIntegrator : public PathIntegrator {            // "-gp"
public:
    GpuSceneBridge gpu;
    void init(char* filename, Parameters& para) { PathIntegrator::init(filename, para); gpu.upload(scene); }
    void render() {
        std::vector<float> out((size_t)width * height * 3);
        wrt_pt_params p = { width, height, samplesPerPixel, maxTracingDepth, /*seed*/ 0, 0, 1, 0.f };
        if (wrt_render_pt(gpu.dev, &gpu.cam, &p, out.data()) != WRT_OK) { fprintf(stderr, "wrt: %s\n", wrt_last_error()); return; }
        for (int i = 0; i < height; i++) for (int j = 0; j < width; j++) {
            const float* c = &out[3 * ((size_t)i * width + j)];
            film->color[i][j] = Color3(c[0], c[1], c[2]);      // already scaled by 1/spp (surfaceIntegrator.cpp:45)
        }
    }
};

class GpuWhittedIntegrator : public WhittedIntegrator {      // "-gr"  (R/src/surfaceIntegrator/whitted.h)
public:
    GpuSceneBridge gpu;
    void init(char* filename, Parameters& para) { WhittedIntegrator::init(filename, para); gpu.upload(scene); }
    void render() {
        std::vector<float> out((size_t)width * height * 3);
        wrt_pt_params p = { width, height, samplesPerPixel, maxTracingDepth, /*seed*/ 0, 0, 1, 0.f };
        if (wrt_render_whitted(gpu.dev, &gpu.cam, &p, out.data()) != WRT_OK) { fprintf(stderr, "wrt: %s\n", wrt_last_error()); return; }
        for (int i = 0; i < height; i++) for (int j = 0; j < width; j++) {
            const float* c = &out[3 * ((size_t)i * width + j)];
            film->color[i][j] = Color3(c[0], c[1], c[2]);      // NaN where the reference's own render() leaves NaN
        }
    }
};

class GpuBidirPathTracing : public BidirPathTracing {        // "-gbpt"
public:
    GpuSceneBridge gpu;
    void init(char* filename, Parameters& para) { BidirPathTracing::init(filename, para); gpu.upload(scene); }
    void render() {
        std::vector<float> out((size_t)width * height * 3);
        // film_scale 1: keep the raw accumulator; BidirPathTracing::outputImage divides by `iterations` and transposes
        wrt_bdpt_params p = { width, height, iterations, minPathLength, maxPathLength, controlLength, 0, 0, 1, 1.f, 0 };
        if (wrt_render_bdpt(gpu.dev, &gpu.cam, &p, out.data()) != WRT_OK) { fprintf(stderr, "wrt: %s\n", wrt_last_error()); return; }
        for (int i = 0; i < height; i++) for (int j = 0; j < width; j++) {
            const float* c = &out[3 * ((size_t)i * width + j)];
            film->color[i][j] = Color3(c[0], c[1], c[2]);
        }
    }
};
