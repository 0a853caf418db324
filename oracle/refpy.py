"""TEST INFRASTRUCTURE ONLY — ctypes front-end for oracle/_ref/libwrt_ref.so.

The library is the UNMODIFIED reference renderer (compiled by oracle/Makefile from
/root/reference) behind the small C harness in oracle/ref_harness.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libwrt_ref.so")
HOOKS_SO = os.path.join(_HERE, "_ref", "libwrt_ref_hooks.so")
REF_ROOT = "/root/reference/Winmad-s-raytracer-v1.0"

_lib = None
_hooks = None


def available():
    return os.path.exists(REF_SO)


def lib():
    """Load the reference library (its static initialisers create debug_bpt.txt in cwd, so load
    it from a scratch directory)."""
    global _lib
    if _lib is None:
        cwd = os.getcwd()
        scratch = tempfile.mkdtemp(prefix="wrt_ref_")
        os.chdir(scratch)
        global _hooks
        try:
            # the interposer first, globally visible: libwrt_ref.so's PLT calls to generateLightSample /
            # generateCameraSample then bind to it (ref_hooks.cpp); lazy binding lets its own RNG references wait
            if os.path.exists(HOOKS_SO):
                _hooks = C.CDLL(HOOKS_SO, mode=C.RTLD_GLOBAL | os.RTLD_LAZY)
            _lib = C.CDLL(REF_SO)
        finally:
            os.chdir(cwd)
        _lib.ref_create.restype = C.c_void_p
        _lib.ref_traverse_calls.restype = C.c_ulonglong
        if _hooks is not None:
            real = [C.cast(getattr(_lib, n), C.c_void_p) for n in
                    ("_ZN16BidirPathTracing19generateLightSampleER14BidirPathState",
                     "_ZN16BidirPathTracing20generateCameraSampleEiR14BidirPathState")]
            _hooks.ref_hooks_set_real(real[0], real[1])
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def fixed_torus_scene(dst_dir=None):
    """torus.scene with its absolute Windows OBJ paths rewritten to the bundled ObjFiles
    (SURVEY.md §8c).  torus_mirror.obj does not exist and silently contributes no geometry."""
    import re
    src = open(os.path.join(REF_ROOT, "torus.scene")).read()
    out = re.sub(r'path="[^"]*\\\\(torus_[a-z]+\.obj)"',
                 lambda m: 'path="%s/ObjFiles/%s"' % (REF_ROOT, m.group(1)), src)
    dst_dir = dst_dir or tempfile.mkdtemp(prefix="wrt_scene_")
    p = os.path.join(dst_dir, "torus_fixed.scene")
    open(p, "w").write(out)
    return p


class RefScene:
    """One reference integrator (+ its Scene).  kind: 'pt', 'bdpt' or 'whitted'."""

    def __init__(self, kind="pt"):
        self.L = lib()
        self.kind = kind
        self.h = C.c_void_p(self.L.ref_create({"pt": 0, "bdpt": 1, "whitted": 2}[kind]))
        self.width = self.height = 0

    # -- construction -------------------------------------------------------------------------
    def load_file(self, path, width, height):
        self.width, self.height = width, height
        return self.L.ref_load_scene_file(self.h, path.encode(), width, height)

    def build(self, materials, kind, data9, matid, lights12, cam12, width, height):
        materials = np.ascontiguousarray(materials, np.float32).reshape(-1, 11)
        kind = np.ascontiguousarray(kind, np.int32)
        data9 = np.ascontiguousarray(data9, np.float32).reshape(-1, 9)
        matid = np.ascontiguousarray(matid, np.int32)
        lights12 = np.ascontiguousarray(lights12, np.float32).reshape(-1, 12)
        cam12 = np.ascontiguousarray(cam12, np.float32)
        self.width, self.height = width, height
        return self.L.ref_build_scene(self.h, len(materials), _fp(materials), len(kind), _ip(kind),
                                      _fp(data9), _ip(matid), len(lights12), _fp(lights12),
                                      _fp(cam12), width, height)

    # -- getters ------------------------------------------------------------------------------
    def prims(self):
        n = self.L.ref_num_prims(self.h)
        kind = np.zeros(n, np.int32); data = np.zeros((n, 9), np.float32); mat = np.zeros(n, np.int32)
        self.L.ref_get_prims(self.h, _ip(kind), _fp(data), _ip(mat))
        return kind, data, mat

    def prim_boxes(self):
        n = self.L.ref_num_prims(self.h)
        b = np.zeros((n, 6), np.float32)
        self.L.ref_get_prim_boxes(self.h, _fp(b))
        return b

    def materials(self):
        n = self.L.ref_num_materials(self.h)
        m = np.zeros((n, 11), np.float32)
        self.L.ref_get_materials(self.h, _fp(m))
        return m

    def lights(self):
        n = self.L.ref_num_lights(self.h)
        l = np.zeros((n, 22), np.float32)
        self.L.ref_get_lights(self.h, _fp(l))
        return l

    def camera(self):
        c = np.zeros(45, np.float32)
        self.L.ref_get_camera(self.h, _fp(c))
        return c

    def scene_sphere(self):
        s = np.zeros(5, np.float32)
        self.L.ref_get_scene_sphere(self.h, _fp(s))
        return s

    def tree(self):
        nn = C.c_longlong(); nr = C.c_longlong(); dep = C.c_int(); dmax = C.c_int()
        self.L.ref_tree_stats(self.h, C.byref(nn), C.byref(nr), C.byref(dep), C.byref(dmax))
        nn, nr = nn.value, nr.value
        t = dict(axis=np.zeros(nn, np.int32), split=np.zeros(nn, np.float32),
                 left=np.zeros(nn, np.int32), right=np.zeros(nn, np.int32),
                 first_ref=np.zeros(nn, np.int32), nref=np.zeros(nn, np.int32),
                 box=np.zeros((nn, 6), np.float32), refs=np.zeros(max(nr, 1), np.int32))
        self.L.ref_tree_flatten(self.h, _ip(t["axis"]), _fp(t["split"]), _ip(t["left"]),
                                _ip(t["right"]), _ip(t["first_ref"]), _ip(t["nref"]),
                                _fp(t["box"]), _ip(t["refs"]))
        t["refs"] = t["refs"][:nr]
        t["depth"] = dep.value; t["depmax"] = dmax.value
        return t

    # -- queries ------------------------------------------------------------------------------
    def intersect(self, rays8, full=False):
        rays8 = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        n = len(rays8)
        prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        if not full:
            self.L.ref_intersect(self.h, _fp(rays8), C.c_longlong(n), _ip(prim), _fp(t),
                                 None, None, None, None)
            return prim, t
        p = np.zeros((n, 3), np.float32); nn = np.zeros((n, 3), np.float32)
        ins = np.zeros(n, np.int32); mat = np.zeros(n, np.int32)
        self.L.ref_intersect(self.h, _fp(rays8), C.c_longlong(n), _ip(prim), _fp(t), _fp(p), _fp(nn),
                             _ip(ins), _ip(mat))
        return prim, t, p, nn, ins, mat

    def occluded(self, q9):
        q9 = np.ascontiguousarray(q9, np.float32).reshape(-1, 9)
        occ = np.zeros(len(q9), np.uint8)
        self.L.ref_occluded(self.h, _fp(q9), C.c_longlong(len(q9)),
                            occ.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return occ

    def generate_rays(self, xy):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        r = np.zeros((len(xy), 8), np.float32)
        self.L.ref_generate_rays(self.h, _fp(xy), C.c_longlong(len(xy)), _fp(r))
        return r

    def render_pt(self, spp, max_depth, seed=5489):
        f = np.zeros((self.height, self.width, 3), np.float32)
        assert self.L.ref_render_pt(self.h, spp, max_depth, C.c_uint(seed), _fp(f)) == 0
        return f

    def render_pt_rows(self, spp, max_depth, seed, row0, row1, col0, col1, want_film=True, row_stride=1):
        """Rows row0, row0+stride, ... < row1, columns [col0, col1): same per-sample body as render()."""
        f = np.zeros((self.height, self.width, 3), np.float32) if want_film else None
        assert self.L.ref_render_pt_rows(self.h, spp, max_depth, C.c_uint(seed), row0, row1, row_stride, col0,
                                         col1, _fp(f) if want_film else None) == 0
        return f

    def render_whitted(self, spp, max_depth, seed=5489):
        f = np.zeros((self.height, self.width, 3), np.float32)
        assert self.L.ref_render_whitted(self.h, spp, max_depth, C.c_uint(seed), _fp(f)) == 0
        return f

    def render_bdpt(self, iterations, seed=5489, control_length=3, max_path_length=10):
        f = np.zeros((self.height, self.width, 3), np.float32)
        assert self.L.ref_render_bdpt(self.h, iterations, C.c_uint(seed), control_length,
                                      max_path_length, _fp(f)) == 0
        return f

    # -- shading known-answer batches, shadow / any-hit queries, RNG tapes (ref_harness.cpp) ------------
    SHADING_IN = {0: 10, 1: 10, 2: 10, 3: 7, 4: 7, 5: 4, 6: 2, 7: 13, 8: 5}
    SHADING_OUT = {0: 9, 1: 9, 2: 2, 3: 10, 4: 12, 5: 5, 6: 1, 7: 13, 8: 8}

    def shading(self, what, inputs, iparam=0):
        a = np.ascontiguousarray(inputs, np.float32).reshape(-1, self.SHADING_IN[what])
        out = np.zeros((len(a), self.SHADING_OUT[what]), np.float32)
        assert self.L.ref_shading_batch(self.h, int(what), int(iparam), _fp(a), C.c_longlong(len(a)), _fp(out)) == 0
        return out

    def shadow_test(self, rays8, target3):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        p = np.ascontiguousarray(target3, np.float32).reshape(-1, 3)
        vis = np.zeros(len(r), np.float32)
        self.L.ref_shadow_test(self.h, _fp(r), _fp(p), C.c_longlong(len(r)), _fp(vis))
        return vis

    def intersect_any(self, rays8):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        hit = np.zeros(len(r), np.uint8)
        self.L.ref_intersect_any(self.h, _fp(r), C.c_longlong(len(r)), hit.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return hit

    def render_pt_tape(self, spp, max_depth, seed=5489, stride=96):
        """The unmodified per-sample code with a recorder around every sample: returns (film, tape[n][stride],
        sample_rgb[n][3], draws[n]) with n = H*W*spp in the render loop's order (pixel-major, then sample)."""
        n = self.height * self.width * spp
        tape = np.zeros((n, stride), np.float32); rgb = np.zeros((n, 3), np.float32); draws = np.zeros(n, np.int32)
        film = np.zeros((self.height, self.width, 3), np.float32)
        assert self.L.ref_render_pt_tape(self.h, spp, max_depth, C.c_uint(seed), stride, _fp(tape), _fp(rgb), _ip(draws), _fp(film)) == 0
        assert draws.max() <= stride, "tape stride too short: a sample drew %d numbers" % draws.max()
        return film, tape, rgb, draws

    def render_bdpt_tape(self, iterations, seed=5489, stride=160, control_length=3, max_path_length=10):
        """BidirPathTracing::render() with the interposed recorder (ref_hooks.cpp): returns (raw film, tape, draws) where
        tape[(it*W*H + p)*2 + {0: light, 1: camera}][stride] holds the stream at the start of that path and draws the
        numbers each path consumed (the last camera path's count is unknown: -1)."""
        lib()
        assert _hooks is not None, "oracle/_ref/libwrt_ref_hooks.so not built"
        paths = self.width * self.height * iterations
        tape = np.zeros((2 * paths, stride), np.float32)
        pos = np.zeros(2 * paths, np.int64)
        _hooks.ref_hooks_start(_fp(tape), stride, C.c_longlong(paths), pos.ctypes.data_as(C.POINTER(C.c_longlong)))
        try:
            film = self.render_bdpt(iterations, seed=seed, control_length=control_length, max_path_length=max_path_length)
        finally:
            lc, cc = C.c_longlong(), C.c_longlong()
            _hooks.ref_hooks_stop(C.byref(lc), C.byref(cc))
        assert lc.value == paths and cc.value == paths, "hooks saw %d light / %d camera paths, expected %d" % (lc.value, cc.value, paths)
        # order of execution: per iteration all light paths, then all camera paths
        npix = self.width * self.height
        order = np.concatenate([np.concatenate([(np.arange(npix) + it * npix) * 2, (np.arange(npix) + it * npix) * 2 + 1]) for it in range(iterations)])
        draws = np.full(2 * paths, -1, np.int64)
        draws[order[:-1]] = pos[order[1:]] - pos[order[:-1]]
        assert draws.max() <= stride, "tape stride too short: a path drew %d numbers" % draws.max()
        return film, tape, draws

    def traverse_calls(self):
        return int(self.L.ref_traverse_calls())

    def reset_traverse_calls(self):
        self.L.ref_reset_traverse_calls()


def make_rays(od6):
    od6 = np.ascontiguousarray(od6, np.float32).reshape(-1, 6)
    r = np.zeros((len(od6), 8), np.float32)
    lib().ref_make_rays(_fp(od6), C.c_longlong(len(od6)), _fp(r))
    return r
