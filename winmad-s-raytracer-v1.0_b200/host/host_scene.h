// Host-side scene: what the reference's Scene::init leaves behind (R/src/scene/scene.cpp:469-489),
// held as flat arrays ready to cross the C ABI (include/wrt.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/wrt.h"

namespace wrt {

// The reference's numeric kernel (R/src/math/math.h:13-18, math.cpp:8-11).
typedef float Real;
static const Real kEps = 1e-3f;
static const Real kInf = 1e7f;
static inline int cmp_eps(Real x) { return (x < -kEps) ? -1 : (x > kEps); }

struct FlatTree {
    std::vector<int32_t> axis, left, right, first_ref, n_ref;
    std::vector<float> split;
    std::vector<int32_t> refs;
    float root_box[6];
    int dep_max = 0;
    int depth = 0;
    void clear() {
        axis.clear(); left.clear(); right.clear(); first_ref.clear(); n_ref.clear();
        split.clear(); refs.clear(); depth = 0;
    }
};

struct HostScene {
    // Scene::objs order
    std::vector<int32_t> prim_kind;
    std::vector<float> prim_data;  // 9 per prim
    std::vector<int32_t> prim_matid;
    std::vector<float> materials;  // 11 per material
    std::vector<float> lights;     // 12 per light
    bool has_camera = false;
    float cam_args[12];            // pos fwd up xres yres fov (as given to Camera::setup)
    wrt_camera camera;
    FlatTree tree;
    bool tree_built = false;
    float scene_sphere[5];
    int n_prims() const { return (int)prim_kind.size(); }
};

// kd_build.cpp
void prim_box(int kind, const float* d9, float box6[6]);  // Triangle::setBox / Sphere::setBox + AABB::extend
bool build_kdtree(HostScene& hs, std::string& err);

// scene_io.cpp
bool load_scene_file(const char* path, HostScene& hs, std::string& err);
bool load_obj_triangles(const char* path, std::vector<std::vector<float> >& shapes_tris,
                        std::vector<std::string>& shape_names);
void camera_setup(const float pos[3], const float fwd[3], const float up[3], float xres, float yres,
                  float fov, wrt_camera* out);
void camera_generate_ray(const wrt_camera& cam, float x, float y, wrt_ray* out);
void make_ray(const float* od6, wrt_ray* out);
bool film_write(const char* path, const float* film, int w, int h, float scale, float gamma,
                std::string& err);
bool save_cache(const HostScene& hs, const char* path, std::string& err);
bool load_cache(const char* path, HostScene& hs, std::string& err);

void set_error(const std::string& s);

}  // namespace wrt
