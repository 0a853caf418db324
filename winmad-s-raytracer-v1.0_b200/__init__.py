"""B200-native render core for Winmad's raytracer — Python host side over the C ABI (include/wrt.h).

The classes mirror the reference's own interface for this path
(R = /root/reference/Winmad-s-raytracer-v1.0):

    Parameters          R/src/parameters.{h,cpp}         load_parameters(file)
    HostScene           R/src/scene/scene.cpp:230-489    Scene::init: load .scene/OBJ, build the KD-tree
    Scene               R/src/scene/scene.h:44-50        intersect / shadowRayTest / occluded (batched)
    PathIntegrator      R/src/surfaceIntegrator/pathIntegrator.{h,cpp}      init / render / outputImage
    BidirPathTracing    R/src/surfaceIntegrator/bidirPathTracing.{h,cpp}    init / render / outputImage

Everything that computes runs in libwrt_b200.so (hand-written CUDA for sm_100a).  There is no CPU
fallback: if the library is missing or no CUDA device is usable, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("WRT_B200_LIB", "libwrt_b200.so"))   # override: A/B builds made by tools/build_variant.sh

OK = 0
TRAVERSE_EXACT, TRAVERSE_PRUNED = 0, 1
PRIM_TRIANGLE, PRIM_SPHERE = 0, 1
EPS = np.float32(1e-3)
INF = np.float32(1e7)

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


class WrtError(RuntimeError):
    pass


class KdTree(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("axis", _i32p), ("split", _f32p), ("left", _i32p), ("right", _i32p),
                ("first_ref", _i32p), ("n_ref", _i32p), ("n_refs", C.c_int64), ("refs", _i32p),
                ("root_box", C.c_float * 6)]


class SceneDesc(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("prim_kind", _i32p), ("prim_data", _f32p), ("prim_matid", _i32p),
                ("n_materials", C.c_int32), ("materials", _f32p), ("n_lights", C.c_int32), ("lights", _f32p),
                ("tree", KdTree)]


class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("forward", C.c_float * 3), ("image_plane_dist", C.c_float),
                ("x_res", C.c_float), ("y_res", C.c_float), ("raster_to_world", C.c_float * 16),
                ("world_to_raster", C.c_float * 16)]

    @classmethod
    def from_ref_array(cls, c45):
        """Build from the oracle's camera dump (oracle/ref_harness.cpp ref_get_camera)."""
        c = cls()
        c.pos[:] = [float(v) for v in c45[0:3]]
        c.forward[:] = [float(v) for v in c45[3:6]]
        c.x_res, c.y_res, c.image_plane_dist = float(c45[9]), float(c45[10]), float(c45[12])
        c.raster_to_world[:] = [float(v) for v in c45[13:29]]
        c.world_to_raster[:] = [float(v) for v in c45[29:45]]
        return c


class PtParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("seed", C.c_uint32), ("sample_first", C.c_int32), ("sample_stride", C.c_int32),
                ("film_scale", C.c_float)]


class BdptParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("iterations", C.c_int32),
                ("min_path_length", C.c_int32), ("max_path_length", C.c_int32), ("control_length", C.c_int32),
                ("seed", C.c_uint32), ("iter_first", C.c_int32), ("iter_stride", C.c_int32),
                ("film_scale", C.c_float), ("transpose_output", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("samples", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("inner_visits", C.c_uint64), ("leaf_visits", C.c_uint64),
                ("tri_tests", C.c_uint64), ("sphere_tests", C.c_uint64), ("last_render_ms", C.c_double),
                ("last_trace_ms", C.c_double), ("extend_ms", C.c_double), ("shade_ms", C.c_double),
                ("shadow_ms", C.c_double), ("extend_launches", C.c_uint64), ("extend_rays", C.c_uint64),
                ("reduce_ms", C.c_double), ("devices_used", C.c_uint64)]


EXPORTS = [
    "wrt_version", "wrt_last_error", "wrt_device_count", "wrt_set_device", "wrt_init", "wrt_shutdown",
    "wrt_host_scene_load", "wrt_host_scene_from_arrays", "wrt_host_scene_build_kdtree", "wrt_host_scene_desc",
    "wrt_host_scene_camera", "wrt_host_scene_sphere", "wrt_host_scene_free", "wrt_host_scene_save",
    "wrt_host_scene_load_cache", "wrt_camera_setup", "wrt_camera_generate_rays", "wrt_make_rays",
    "wrt_film_write", "wrt_scene_create", "wrt_scene_destroy", "wrt_scene_set_traversal", "wrt_scene_set_counting",
    "wrt_get_stats",
    "wrt_reset_stats", "wrt_trace_closest", "wrt_trace_closest_full", "wrt_trace_any", "wrt_trace_shadow",
    "wrt_trace_occluded", "wrt_trace_closest_dev", "wrt_trace_occluded_dev", "wrt_trace_count_visits",
    "wrt_render_pt", "wrt_render_pt_dev", "wrt_render_whitted", "wrt_render_whitted_dev", "wrt_render_bdpt", "wrt_render_bdpt_dev", "wrt_film_resolve_dev",
    "wrt_debug_shading", "wrt_debug_set_rng_tape",
]

_lib = None


def lib():
    """Load libwrt_b200.so.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WrtError("libwrt_b200.so is not built (run `python __graft_entry__.py` or "
                           "`python winmad-s-raytracer-v1.0_b200/build.py`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.wrt_version.restype = C.c_char_p
        L.wrt_last_error.restype = C.c_char_p
        L.wrt_scene_destroy.restype = None
        L.wrt_host_scene_free.restype = None
        _lib = L
    return _lib


def _check(rc, what=""):
    if rc != OK:
        raise WrtError("%s failed (code %d): %s" % (what, rc, lib().wrt_last_error().decode(errors="replace")))


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


def device_count():
    n = C.c_int(0)
    rc = lib().wrt_device_count(C.byref(n))
    return n.value if rc == OK else 0


def set_device(ordinal):
    _check(lib().wrt_set_device(int(ordinal)), "wrt_set_device")


def init(n_gpus=0, device_ids=None):
    """wrt_init: scenes created afterwards are replicated on these devices; one render call drives them all."""
    ids = None
    if device_ids is not None:
        arr = (C.c_int * len(device_ids))(*[int(d) for d in device_ids])
        ids, n_gpus = arr, len(device_ids)
    _check(lib().wrt_init(int(n_gpus), ids), "wrt_init")


def shutdown():
    _check(lib().wrt_shutdown(), "wrt_shutdown")


def make_rays(origin_dir6):
    """Ray(origin, dir) constructor (R/src/geometry/ray.h:14-16) on a batch: (n,6) -> (n,8)."""
    od = _f32(origin_dir6, (-1, 6))
    out = np.zeros((len(od), 8), np.float32)
    _check(lib().wrt_make_rays(_ptr(od, _f32p), C.c_size_t(len(od)), out.ctypes.data_as(C.c_void_p)), "wrt_make_rays")
    return out


def camera_setup(pos, forward, up, x_res, y_res, fov):
    cam = Camera()
    p, f, u = _f32(pos), _f32(forward), _f32(up)
    _check(lib().wrt_camera_setup(_ptr(p, _f32p), _ptr(f, _f32p), _ptr(u, _f32p), C.c_float(x_res),
                                  C.c_float(y_res), C.c_float(fov), C.byref(cam)), "wrt_camera_setup")
    return cam


def generate_rays(cam, xy):
    """Camera::generateRay on a batch of raster positions: (n,2) -> (n,8)."""
    xy = _f32(xy, (-1, 2))
    out = np.zeros((len(xy), 8), np.float32)
    _check(lib().wrt_camera_generate_rays(C.byref(cam), _ptr(xy, _f32p), C.c_size_t(len(xy)),
                                          out.ctypes.data_as(C.c_void_p)), "wrt_camera_generate_rays")
    return out


def film_write(path, film, scale=1.0, gamma=2.2):
    film = _f32(film)
    h, w = film.shape[0], film.shape[1]
    _check(lib().wrt_film_write(path.encode(), _ptr(film, _f32p), w, h, C.c_float(scale), C.c_float(gamma)),
           "wrt_film_write")


def film_resolve_dev(d_film, width, height, scale, gamma, d_rgb, stream=None):
    """ImageFilm::outputImage's scale/clamp/gamma/8-bit on device buffers (raw pointers)."""
    _check(lib().wrt_film_resolve_dev(C.c_void_p(d_film), int(width), int(height), C.c_float(scale), C.c_float(gamma),
                                      C.c_void_p(d_rgb), C.c_void_p(stream or 0)), "wrt_film_resolve_dev")


class Parameters:
    """R/src/parameters.{h,cpp}: eight positional ints; '#' lines are comments."""
    FIELDS = ["MAX_TRACING_DEPTH", "SAMPLES_PER_PIXEL", "SAMPLES_OF_LIGHT", "SAMPLES_OF_HEMISPHERE",
              "WIDTH", "HEIGHT", "PHONG_POWER_INDEX", "POINT_LIGHT_NUM"]

    def __init__(self, **kw):
        self.MAX_TRACING_DEPTH, self.SAMPLES_PER_PIXEL = 7, 1
        self.SAMPLES_OF_LIGHT, self.SAMPLES_OF_HEMISPHERE = 8, 4
        self.WIDTH, self.HEIGHT = 512, 512
        self.PHONG_POWER_INDEX, self.POINT_LIGHT_NUM = 5, 400
        for k, v in kw.items():
            setattr(self, k, v)

    def load_parameters(self, filename):
        vals = []
        with open(filename) as f:
            for tok in f.read().split():
                if tok.startswith("#"):
                    continue
                try:
                    vals.append(int(tok))  # atoi
                except ValueError:
                    vals.append(0)
        for name, v in zip(self.FIELDS, vals + [-1] * 8):
            setattr(self, name, v)
        return self


class HostScene:
    """What Scene::init leaves behind, on the host: primitives, materials, lights, camera, KD-tree."""

    def __init__(self, handle):
        self._h = handle
        self._keep = None

    @classmethod
    def load(cls, scene_file, build=True):
        h = C.c_void_p()
        _check(lib().wrt_host_scene_load(os.fsencode(scene_file), C.byref(h)), "wrt_host_scene_load")
        hs = cls(h)
        if build:
            hs.build_kdtree()
        return hs

    @classmethod
    def from_arrays(cls, materials, prim_kind, prim_data, prim_matid, lights, cam12=None, build=True):
        m = _f32(materials, (-1, 11)); k = np.ascontiguousarray(prim_kind, np.int32)
        d = _f32(prim_data, (-1, 9)); mi = np.ascontiguousarray(prim_matid, np.int32)
        l = _f32(lights, (-1, 12))
        cam = _f32(cam12) if cam12 is not None else None
        h = C.c_void_p()
        _check(lib().wrt_host_scene_from_arrays(len(m), _ptr(m, _f32p), len(k), _ptr(k, _i32p), _ptr(d, _f32p),
                                                _ptr(mi, _i32p), len(l), _ptr(l, _f32p),
                                                _ptr(cam, _f32p) if cam is not None else None, C.byref(h)),
               "wrt_host_scene_from_arrays")
        hs = cls(h)
        if build:
            hs.build_kdtree()
        return hs

    @classmethod
    def load_cache(cls, path):
        h = C.c_void_p()
        _check(lib().wrt_host_scene_load_cache(os.fsencode(path), C.byref(h)), "wrt_host_scene_load_cache")
        return cls(h)

    def save(self, path):
        _check(lib().wrt_host_scene_save(self._h, os.fsencode(path)), "wrt_host_scene_save")

    def build_kdtree(self):
        _check(lib().wrt_host_scene_build_kdtree(self._h), "wrt_host_scene_build_kdtree")

    def desc(self):
        d = SceneDesc()
        _check(lib().wrt_host_scene_desc(self._h, C.byref(d)), "wrt_host_scene_desc")
        return d

    def camera(self):
        c = Camera()
        _check(lib().wrt_host_scene_camera(self._h, C.byref(c)), "wrt_host_scene_camera")
        return c

    def scene_sphere(self):
        s = (C.c_float * 5)()
        _check(lib().wrt_host_scene_sphere(self._h, s), "wrt_host_scene_sphere")
        return np.array(s[:], np.float32)

    # numpy copies, for tests and tooling
    def arrays(self):
        d = self.desc()
        n, T = d.n_prims, d.tree
        out = dict(
            prim_kind=np.ctypeslib.as_array(d.prim_kind, (n,)).copy() if n else np.zeros(0, np.int32),
            prim_data=np.ctypeslib.as_array(d.prim_data, (n, 9)).copy() if n else np.zeros((0, 9), np.float32),
            prim_matid=np.ctypeslib.as_array(d.prim_matid, (n,)).copy() if n else np.zeros(0, np.int32),
            materials=np.ctypeslib.as_array(d.materials, (d.n_materials, 11)).copy() if d.n_materials else np.zeros((0, 11), np.float32),
            lights=np.ctypeslib.as_array(d.lights, (d.n_lights, 12)).copy() if d.n_lights else np.zeros((0, 12), np.float32),
        )
        if T.n_nodes:
            nn = T.n_nodes
            out["tree"] = dict(
                axis=np.ctypeslib.as_array(T.axis, (nn,)).copy(), split=np.ctypeslib.as_array(T.split, (nn,)).copy(),
                left=np.ctypeslib.as_array(T.left, (nn,)).copy(), right=np.ctypeslib.as_array(T.right, (nn,)).copy(),
                first_ref=np.ctypeslib.as_array(T.first_ref, (nn,)).copy(),
                nref=np.ctypeslib.as_array(T.n_ref, (nn,)).copy(),
                refs=np.ctypeslib.as_array(T.refs, (T.n_refs,)).copy() if T.n_refs else np.zeros(0, np.int32),
                root_box=np.array(T.root_box[:], np.float32))
        return out

    def close(self):
        if self._h:
            lib().wrt_host_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def desc_from_arrays(prim_kind, prim_data, prim_matid, materials, lights, tree):
    """Build a wrt_scene_desc from numpy arrays (e.g. a tree flattened out of the reference itself).
    Returns (desc, keepalive)."""
    keep = dict(
        kind=np.ascontiguousarray(prim_kind, np.int32), data=_f32(prim_data, (-1, 9)),
        mat=np.ascontiguousarray(prim_matid, np.int32), materials=_f32(materials, (-1, 11)),
        lights=_f32(lights, (-1, 12)),
        axis=np.ascontiguousarray(tree["axis"], np.int32), split=_f32(tree["split"]),
        left=np.ascontiguousarray(tree["left"], np.int32), right=np.ascontiguousarray(tree["right"], np.int32),
        first_ref=np.ascontiguousarray(tree["first_ref"], np.int32), nref=np.ascontiguousarray(tree["nref"], np.int32),
        refs=np.ascontiguousarray(tree["refs"], np.int32))
    d = SceneDesc()
    d.n_prims = len(keep["kind"]); d.prim_kind = _ptr(keep["kind"], _i32p); d.prim_data = _ptr(keep["data"], _f32p)
    d.prim_matid = _ptr(keep["mat"], _i32p)
    d.n_materials = len(keep["materials"]); d.materials = _ptr(keep["materials"], _f32p)
    d.n_lights = len(keep["lights"]); d.lights = _ptr(keep["lights"], _f32p)
    T = d.tree
    T.n_nodes = len(keep["axis"]); T.axis = _ptr(keep["axis"], _i32p); T.split = _ptr(keep["split"], _f32p)
    T.left = _ptr(keep["left"], _i32p); T.right = _ptr(keep["right"], _i32p)
    T.first_ref = _ptr(keep["first_ref"], _i32p); T.n_ref = _ptr(keep["nref"], _i32p)
    T.n_refs = len(keep["refs"]); T.refs = _ptr(keep["refs"], _i32p)
    rb = tree["root_box"] if "root_box" in tree else tree["box"][0]
    T.root_box[:] = [float(v) for v in np.asarray(rb).ravel()[:6]]
    return d, keep


class Scene:
    """Device-resident scene with the reference's query methods (R/src/scene/scene.h:44-50), batched."""

    def __init__(self, source, keepalive=None):
        self._sc = C.c_void_p()
        self._keep = keepalive
        self.host = None
        if isinstance(source, HostScene):
            self.host = source
            d = source.desc()
        else:
            d = source
        _check(lib().wrt_scene_create(C.byref(d), C.byref(self._sc)), "wrt_scene_create")

    def set_traversal(self, mode):
        _check(lib().wrt_scene_set_traversal(self._sc, int(mode)), "wrt_scene_set_traversal")

    def set_counting(self, on):
        _check(lib().wrt_scene_set_counting(self._sc, int(on)), "wrt_scene_set_counting")

    def stats(self):
        s = Stats()
        _check(lib().wrt_get_stats(self._sc, C.byref(s)), "wrt_get_stats")
        return s

    def reset_stats(self):
        _check(lib().wrt_reset_stats(self._sc), "wrt_reset_stats")

    # Geometry* Scene::intersect(const Ray&, Intersection&)
    def intersect(self, rays8, full=False):
        r = _f32(rays8, (-1, 8)); n = len(r)
        prim = np.full(n, -1, np.int32); t = np.zeros(n, np.float32)
        if not full:
            _check(lib().wrt_trace_closest(self._sc, r.ctypes.data_as(C.c_void_p), C.c_size_t(n), _ptr(prim, _i32p),
                                           _ptr(t, _f32p)), "wrt_trace_closest")
            return prim, t
        p = np.zeros((n, 3), np.float32); nn = np.zeros((n, 3), np.float32)
        ins = np.zeros(n, np.int32); mat = np.zeros(n, np.int32)
        _check(lib().wrt_trace_closest_full(self._sc, r.ctypes.data_as(C.c_void_p), C.c_size_t(n), _ptr(prim, _i32p),
                                            _ptr(t, _f32p), _ptr(p, _f32p), _ptr(nn, _f32p), _ptr(ins, _i32p),
                                            _ptr(mat, _i32p)), "wrt_trace_closest_full")
        return prim, t, p, nn, ins, mat

    # bool Scene::intersect(const Ray&)
    def intersect_any(self, rays8):
        r = _f32(rays8, (-1, 8)); n = len(r)
        hit = np.zeros(n, np.uint8)
        _check(lib().wrt_trace_any(self._sc, r.ctypes.data_as(C.c_void_p), C.c_size_t(n), _ptr(hit, _u8p)), "wrt_trace_any")
        return hit

    # Real Scene::shadowRayTest(const Ray&, const Vector3& p)
    def shadowRayTest(self, rays8, p3):
        r = _f32(rays8, (-1, 8)); p = _f32(p3, (-1, 3)); n = len(r)
        vis = np.zeros(n, np.float32)
        _check(lib().wrt_trace_shadow(self._sc, r.ctypes.data_as(C.c_void_p), _ptr(p, _f32p), C.c_size_t(n),
                                      _ptr(vis, _f32p)), "wrt_trace_shadow")
        return vis

    # bool Scene::occluded(p1, dir, p2)
    def occluded(self, p1_dir_p2):
        q = _f32(p1_dir_p2, (-1, 9)); n = len(q)
        occ = np.zeros(n, np.uint8)
        _check(lib().wrt_trace_occluded(self._sc, _ptr(q, _f32p), C.c_size_t(n), _ptr(occ, _u8p)), "wrt_trace_occluded")
        return occ

    def count_visits(self, rays8):
        r = _f32(rays8, (-1, 8))
        _check(lib().wrt_trace_count_visits(self._sc, r.ctypes.data_as(C.c_void_p), C.c_size_t(len(r))),
               "wrt_trace_count_visits")
        s = self.stats()
        return dict(inner=s.inner_visits, leaf=s.leaf_visits, tri=s.tri_tests, sphere=s.sphere_tests, rays=len(r))

    # device-pointer variants (torch tensors or raw ints)
    def intersect_dev(self, d_rays, n, d_prim, d_t, stream=None):
        _check(lib().wrt_trace_closest_dev(self._sc, C.c_void_p(d_rays), C.c_size_t(n), C.c_void_p(d_prim),
                                           C.c_void_p(d_t), C.c_void_p(stream or 0)), "wrt_trace_closest_dev")

    def occluded_dev(self, d_q9, n, d_occ, stream=None):
        _check(lib().wrt_trace_occluded_dev(self._sc, C.c_void_p(d_q9), C.c_size_t(n), C.c_void_p(d_occ),
                                            C.c_void_p(stream or 0)), "wrt_trace_occluded_dev")

    def render_pt(self, cam, params, film=None):
        film = np.zeros((params.height, params.width, 3), np.float32) if film is None else film
        _check(lib().wrt_render_pt(self._sc, C.byref(cam), C.byref(params), _ptr(film, _f32p)), "wrt_render_pt")
        return film

    def render_pt_dev(self, cam, params, d_film, stream=None):
        _check(lib().wrt_render_pt_dev(self._sc, C.byref(cam), C.byref(params), C.c_void_p(d_film),
                                       C.c_void_p(stream or 0)), "wrt_render_pt_dev")

    def render_whitted(self, cam, params, film=None):
        """WhittedIntegrator (whitted.cpp:17-113); params: PtParams (max_depth = MAX_TRACING_DEPTH)."""
        film = np.zeros((params.height, params.width, 3), np.float32) if film is None else film
        _check(lib().wrt_render_whitted(self._sc, C.byref(cam), C.byref(params), _ptr(film, _f32p)), "wrt_render_whitted")
        return film

    def render_whitted_dev(self, cam, params, d_film, stream=None):
        _check(lib().wrt_render_whitted_dev(self._sc, C.byref(cam), C.byref(params), C.c_void_p(d_film),
                                            C.c_void_p(stream or 0)), "wrt_render_whitted_dev")

    def render_bdpt(self, cam, params, film=None):
        film = np.zeros((params.height, params.width, 3), np.float32) if film is None else film
        _check(lib().wrt_render_bdpt(self._sc, C.byref(cam), C.byref(params), _ptr(film, _f32p)), "wrt_render_bdpt")
        return film

    def render_bdpt_dev(self, cam, params, d_film, stream=None):
        _check(lib().wrt_render_bdpt_dev(self._sc, C.byref(cam), C.byref(params), C.c_void_p(d_film),
                                         C.c_void_p(stream or 0)), "wrt_render_bdpt_dev")

    # diagnostics (include/wrt.h): shading known-answer evaluation and RNG replay
    SHADING_IN = {0: 10, 1: 10, 2: 10, 3: 7, 4: 7, 5: 4, 6: 2, 7: 13, 8: 5}
    SHADING_OUT = {0: 9, 1: 9, 2: 2, 3: 10, 4: 12, 5: 5, 6: 1, 7: 13, 8: 8}

    def debug_shading(self, what, inputs, iparam=0, cam=None):
        a = _f32(inputs, (-1, self.SHADING_IN[what])); n = len(a)
        out = np.zeros((n, self.SHADING_OUT[what]), np.float32)
        _check(lib().wrt_debug_shading(self._sc, C.byref(cam) if cam is not None else None, int(what), int(iparam),
                                       _ptr(a, _f32p), C.c_size_t(n), _ptr(out, _f32p)), "wrt_debug_shading")
        return out

    def set_rng_tape(self, tape, stride):
        if tape is None:
            _check(lib().wrt_debug_set_rng_tape(self._sc, None, C.c_size_t(0), C.c_uint32(0)), "wrt_debug_set_rng_tape")
            return
        t = _f32(tape).ravel()
        _check(lib().wrt_debug_set_rng_tape(self._sc, _ptr(t, _f32p), C.c_size_t(len(t)), C.c_uint32(int(stride))), "wrt_debug_set_rng_tape")

    def close(self):
        if self._sc:
            lib().wrt_scene_destroy(self._sc)
            self._sc = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SurfaceIntegrator:
    """R/src/surfaceIntegrator/surfaceIntegrator.h:14-34."""

    def __init__(self):
        self.width = self.height = 0
        self.samplesPerPixel = 1
        self.scene = None      # device Scene
        self.host_scene = None
        self.film = None       # numpy H x W x 3, linear radiance (ImageFilm::color)
        self.seed = 0

    def _init_scene(self, scene_file, para):
        self.host_scene = scene_file if isinstance(scene_file, HostScene) else HostScene.load(scene_file)
        self.scene = Scene(self.host_scene)
        self.height, self.width = para.HEIGHT, para.WIDTH
        self.film = np.zeros((self.height, self.width, 3), np.float32)

    def outputImage(self, filename):
        film_write(filename, self.film, 1.0, 2.2)


class PathIntegrator(SurfaceIntegrator):
    """PathIntegrator::init / render / outputImage (pathIntegrator.cpp:3-15, surfaceIntegrator.cpp:14-51)."""

    def init(self, scene_file, para):
        self.maxTracingDepth = para.MAX_TRACING_DEPTH
        self.samplesPerPixel = para.SAMPLES_PER_PIXEL
        self.samplesOfLight = para.SAMPLES_OF_LIGHT
        self.samplesOfHemisphere = para.SAMPLES_OF_HEMISPHERE
        self._init_scene(scene_file, para)
        return self

    def params(self, sample_first=0, sample_stride=1, film_scale=0.0):
        return PtParams(self.width, self.height, self.samplesPerPixel, self.maxTracingDepth, self.seed,
                        sample_first, sample_stride, film_scale)

    def render(self):
        self.film = self.scene.render_pt(self.host_scene.camera(), self.params())
        return self.film


class WhittedIntegrator(SurfaceIntegrator):
    """WhittedIntegrator::init / render / outputImage (whitted.cpp:3-15, surfaceIntegrator.cpp:14-51)."""

    def init(self, scene_file, para):
        self.maxTracingDepth = para.MAX_TRACING_DEPTH
        self.samplesPerPixel = para.SAMPLES_PER_PIXEL
        self._init_scene(scene_file, para)
        return self

    def params(self, sample_first=0, sample_stride=1, film_scale=0.0):
        return PtParams(self.width, self.height, self.samplesPerPixel, self.maxTracingDepth, self.seed,
                        sample_first, sample_stride, film_scale)

    def render(self):
        self.film = self.scene.render_whitted(self.host_scene.camera(), self.params())
        return self.film


class BidirPathTracing(SurfaceIntegrator):
    """BidirPathTracing::init / render / outputImage (bidirPathTracing.cpp:5-46)."""

    def init(self, scene_file, para):
        self.minPathLength, self.maxPathLength, self.iterations = 0, 10, 1
        self.samplesPerPixel = para.SAMPLES_PER_PIXEL
        self._init_scene(scene_file, para)
        self.controlLength = 3
        return self

    def params(self, iter_first=0, iter_stride=1, film_scale=0.0, transpose_output=0):
        return BdptParams(self.width, self.height, self.iterations, self.minPathLength, self.maxPathLength,
                          self.controlLength, self.seed, iter_first, iter_stride, film_scale, transpose_output)

    def render(self):
        # raw accumulator film->color[a][b] scaled by 1/iterations; outputImage transposes (:29-46)
        self.film = self.scene.render_bdpt(self.host_scene.camera(), self.params())
        return self.film

    def outputImage(self, filename):
        film_write(filename, np.ascontiguousarray(self.film.transpose(1, 0, 2)), 1.0, 2.2)


# ---- multi-GPU: one process per GPU, samples sharded, one sum-reduce of the float film --------------
def shard_pt(params, rank, world):
    """Rank `rank` of `world` renders samples k = rank, rank+world, ... of the SAME stratification grid and
    RNG keys (keyed on pixel and global sample index), pre-scaled by 1/spp: the rank films sum to the
    1-GPU image."""
    return PtParams(params.width, params.height, params.spp, params.max_depth, params.seed,
                    params.sample_first + rank * max(params.sample_stride, 1),
                    max(params.sample_stride, 1) * world,
                    params.film_scale if params.film_scale != 0.0 else 1.0 / params.spp)


def shard_bdpt(params, rank, world):
    """Rank `rank` runs iterations it = rank, rank+world, ... (camera path p only ever reads light path p,
    bidirPathTracing.cpp:222-229, so iterations are independent)."""
    return BdptParams(params.width, params.height, params.iterations, params.min_path_length,
                      params.max_path_length, params.control_length, params.seed,
                      params.iter_first + rank * max(params.iter_stride, 1), max(params.iter_stride, 1) * world,
                      params.film_scale if params.film_scale != 0.0 else 1.0 / params.iterations,
                      params.transpose_output)


def reduce_film(film_tensor, dst=0):
    """Sum the per-rank float films onto rank `dst` (NCCL over NVLink on GPUs, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_tensor, dst, op=dist.ReduceOp.SUM)
    return film_tensor
