// C-ABI entry points that run on the host only (include/wrt.h, "host side" block).
#include <cstring>
#include <new>
#include <exception>
#include "host_scene.h"

struct wrt_host_scene {
    wrt::HostScene hs;
};

namespace wrt {
static thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
}  // namespace wrt

using wrt::set_error;

// No C++ exception may cross the C boundary (a corrupt file or an absurd size must not terminate the caller).
#define WRT_GUARD_BEGIN try {
#define WRT_GUARD_END(h_to_free) } catch (const std::bad_alloc&) { delete (h_to_free); set_error("out of memory"); return WRT_ERR_NOMEM; } \
    catch (const std::exception& e__) { delete (h_to_free); set_error(std::string("internal error: ") + e__.what()); return WRT_ERR_INVALID; }

extern "C" {

const char* wrt_version(void) { return "wrt-b200 0.1 (sm_100a)"; }
const char* wrt_last_error(void) { return wrt::g_last_error.c_str(); }

int wrt_host_scene_load(const char* scene_file, wrt_host_scene** out)
{
    if (!scene_file || !out) { set_error("wrt_host_scene_load: null argument"); return WRT_ERR_INVALID; }
    wrt_host_scene* h = new (std::nothrow) wrt_host_scene();
    if (!h) { set_error("out of memory"); return WRT_ERR_NOMEM; }
    WRT_GUARD_BEGIN
    std::string err;
    if (!wrt::load_scene_file(scene_file, h->hs, err)) { set_error(err); delete h; return WRT_ERR_IO; }
    WRT_GUARD_END(h)
    *out = h;
    return WRT_OK;
}

int wrt_host_scene_from_arrays(int32_t n_materials, const float* materials11, int32_t n_prims,
                               const int32_t* prim_kind, const float* prim_data9, const int32_t* prim_matid,
                               int32_t n_lights, const float* lights12, const float* cam12,
                               wrt_host_scene** out)
{
    if (!out || n_prims < 0 || n_materials < 0 || n_lights < 0 ||
        (n_prims > 0 && (!prim_kind || !prim_data9 || !prim_matid)) ||
        (n_materials > 0 && !materials11) || (n_lights > 0 && !lights12)) {
        set_error("wrt_host_scene_from_arrays: bad argument");
        return WRT_ERR_INVALID;
    }
    for (int32_t i = 0; i < n_prims; i++)
        if (prim_kind[i] != WRT_PRIM_TRIANGLE && prim_kind[i] != WRT_PRIM_SPHERE) {
            set_error("wrt_host_scene_from_arrays: unknown primitive kind");
            return WRT_ERR_INVALID;
        }
    wrt_host_scene* h = new (std::nothrow) wrt_host_scene();
    if (!h) { set_error("out of memory"); return WRT_ERR_NOMEM; }
    wrt::HostScene& hs = h->hs;
    WRT_GUARD_BEGIN
    hs.materials.assign(materials11, materials11 + (size_t)11 * n_materials);
    hs.prim_kind.assign(prim_kind, prim_kind + n_prims);
    hs.prim_data.assign(prim_data9, prim_data9 + (size_t)9 * n_prims);
    hs.prim_matid.assign(prim_matid, prim_matid + n_prims);
    hs.lights.assign(lights12, lights12 + (size_t)12 * n_lights);
    if (cam12) {
        memcpy(hs.cam_args, cam12, sizeof hs.cam_args);
        wrt::camera_setup(cam12, cam12 + 3, cam12 + 6, cam12[9], cam12[10], cam12[11], &hs.camera);
        hs.has_camera = true;
    }
    WRT_GUARD_END(h)
    *out = h;
    return WRT_OK;
}

int wrt_host_scene_build_kdtree(wrt_host_scene* h)
{
    if (!h) { set_error("null scene"); return WRT_ERR_INVALID; }
    WRT_GUARD_BEGIN
    std::string err;
    if (!wrt::build_kdtree(h->hs, err)) { set_error(err); return WRT_ERR_INVALID; }
    WRT_GUARD_END((wrt_host_scene*)nullptr)
    return WRT_OK;
}

int wrt_host_scene_desc(const wrt_host_scene* h, wrt_scene_desc* d)
{
    if (!h || !d) { set_error("null argument"); return WRT_ERR_INVALID; }
    const wrt::HostScene& hs = h->hs;
    memset(d, 0, sizeof *d);
    d->n_prims = hs.n_prims();
    d->prim_kind = hs.prim_kind.data(); d->prim_data = hs.prim_data.data(); d->prim_matid = hs.prim_matid.data();
    d->n_materials = (int32_t)(hs.materials.size() / 11); d->materials = hs.materials.data();
    d->n_lights = (int32_t)(hs.lights.size() / 12); d->lights = hs.lights.data();
    if (hs.tree_built) {
        const wrt::FlatTree& t = hs.tree;
        d->tree.n_nodes = (int32_t)t.axis.size();
        d->tree.axis = t.axis.data(); d->tree.split = t.split.data();
        d->tree.left = t.left.data(); d->tree.right = t.right.data();
        d->tree.first_ref = t.first_ref.data(); d->tree.n_ref = t.n_ref.data();
        d->tree.n_refs = (int64_t)t.refs.size(); d->tree.refs = t.refs.data();
        memcpy(d->tree.root_box, t.root_box, sizeof t.root_box);
    }
    return WRT_OK;
}

int wrt_host_scene_camera(const wrt_host_scene* h, wrt_camera* out)
{
    if (!h || !out) { set_error("null argument"); return WRT_ERR_INVALID; }
    if (!h->hs.has_camera) { set_error("scene has no camera"); return WRT_ERR_INVALID; }
    *out = h->hs.camera;
    return WRT_OK;
}

int wrt_host_scene_sphere(const wrt_host_scene* h, float out5[5])
{
    if (!h || !out5 || !h->hs.tree_built) { set_error("scene sphere needs a built kd-tree"); return WRT_ERR_INVALID; }
    memcpy(out5, h->hs.scene_sphere, 5 * sizeof(float));
    return WRT_OK;
}

void wrt_host_scene_free(wrt_host_scene* h) { delete h; }

int wrt_host_scene_save(const wrt_host_scene* h, const char* path)
{
    if (!h || !path) { set_error("null argument"); return WRT_ERR_INVALID; }
    std::string err;
    if (!wrt::save_cache(h->hs, path, err)) { set_error(err); return WRT_ERR_IO; }
    return WRT_OK;
}

int wrt_host_scene_load_cache(const char* path, wrt_host_scene** out)
{
    if (!path || !out) { set_error("null argument"); return WRT_ERR_INVALID; }
    wrt_host_scene* h = new (std::nothrow) wrt_host_scene();
    if (!h) { set_error("out of memory"); return WRT_ERR_NOMEM; }
    WRT_GUARD_BEGIN
    std::string err;
    if (!wrt::load_cache(path, h->hs, err)) { set_error(err); delete h; return WRT_ERR_IO; }
    WRT_GUARD_END(h)
    *out = h;
    return WRT_OK;
}

int wrt_camera_setup(const float pos[3], const float forward[3], const float up[3], float x_res, float y_res,
                     float fov, wrt_camera* out)
{
    if (!pos || !forward || !up || !out) { set_error("null argument"); return WRT_ERR_INVALID; }
    wrt::camera_setup(pos, forward, up, x_res, y_res, fov, out);
    return WRT_OK;
}

int wrt_camera_generate_rays(const wrt_camera* cam, const float* xy, size_t n, wrt_ray* rays)
{
    if (!cam || (n && (!xy || !rays))) { set_error("null argument"); return WRT_ERR_INVALID; }
    for (size_t i = 0; i < n; i++) wrt::camera_generate_ray(*cam, xy[2 * i], xy[2 * i + 1], &rays[i]);
    return WRT_OK;
}

int wrt_make_rays(const float* od6, size_t n, wrt_ray* rays)
{
    if (n && (!od6 || !rays)) { set_error("null argument"); return WRT_ERR_INVALID; }
    for (size_t i = 0; i < n; i++) wrt::make_ray(od6 + 6 * i, &rays[i]);
    return WRT_OK;
}

int wrt_film_write(const char* path, const float* film, int32_t width, int32_t height, float scale, float gamma)
{
    if (!path || !film || width <= 0 || height <= 0) { set_error("bad argument"); return WRT_ERR_INVALID; }
    std::string err;
    if (!wrt::film_write(path, film, width, height, scale, gamma, err)) { set_error(err); return WRT_ERR_IO; }
    return WRT_OK;
}

}  // extern "C"
