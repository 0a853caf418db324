"""ctypes front-end of the TEST-ONLY hostsim library (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
SO = os.path.join(_HERE, "libwrt_hostsim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        sys.path.insert(0, os.path.join(_ROOT, "winmad-s-raytracer-v1.0_b200"))
        import importlib.util
        spec = importlib.util.spec_from_file_location("wrt_build", os.path.join(_ROOT, "winmad-s-raytracer-v1.0_b200", "build.py"))
        b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
        b.build_hostsim()
        _lib = C.CDLL(SO)
        _lib.hs_num_recs.restype = C.c_longlong
    return _lib


class HostSim:
    def __init__(self, desc, keep=None):
        self._keep = keep
        self.h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = lib().hs_scene_create(C.byref(desc), C.byref(self.h), err)
        if rc:
            raise RuntimeError(err.value.decode())

    def num_recs(self):
        return int(lib().hs_num_recs(self.h))

    def layout_digest(self):
        f = lib().hs_layout_digest
        f.restype = C.c_ulonglong
        return int(f(self.h))

    def trace_closest(self, rays8, pruned=True):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        prim = np.zeros(len(r), np.int32); t = np.zeros(len(r), np.float32)
        lib().hs_trace_closest(self.h, r.ctypes.data_as(C.c_void_p), C.c_size_t(len(r)), int(pruned),
                               prim.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
        return prim, t

    def trace_closest_full(self, rays8, pruned=True):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8); n = len(r)
        prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32); p = np.zeros((n, 3), np.float32)
        nn = np.zeros((n, 3), np.float32); ins = np.zeros(n, np.int32); mat = np.zeros(n, np.int32)
        lib().hs_trace_closest_full(self.h, r.ctypes.data_as(C.c_void_p), C.c_size_t(n), int(pruned),
                                    *[a.ctypes.data_as(C.c_void_p) for a in (prim, t, p, nn, ins, mat)])
        return prim, t, p, nn, ins, mat

    def trace_occluded(self, q9, pruned=True):
        q = np.ascontiguousarray(q9, np.float32).reshape(-1, 9)
        occ = np.zeros(len(q), np.uint8)
        lib().hs_trace_occluded(self.h, q.ctypes.data_as(C.c_void_p), C.c_size_t(len(q)), int(pruned),
                                occ.ctypes.data_as(C.c_void_p))
        return occ

    def count_visits(self, rays8, pruned=False):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        out = (C.c_ulonglong * 4)()
        lib().hs_count_visits(self.h, r.ctypes.data_as(C.c_void_p), C.c_size_t(len(r)), int(pruned), out)
        return dict(inner=out[0], leaf=out[1], tri=out[2], sphere=out[3], rays=len(r))

    def render_pt(self, cam, params, pruned=True):
        film = np.zeros((params.height, params.width, 3), np.float32)
        rays = C.c_ulonglong(0)
        lib().hs_render_pt(self.h, C.byref(cam), C.byref(params), int(pruned), film.ctypes.data_as(C.c_void_p),
                           C.byref(rays))
        return film, rays.value

    def render_whitted(self, cam, params, pruned=True):
        film = np.zeros((params.height, params.width, 3), np.float32)
        rays = C.c_ulonglong(0)
        lib().hs_render_whitted(self.h, C.byref(cam), C.byref(params), int(pruned), film.ctypes.data_as(C.c_void_p),
                                C.byref(rays))
        return film, rays.value

    def render_bdpt(self, cam, params, pruned=True):
        film = np.zeros((params.height, params.width, 3), np.float32)
        rays = C.c_ulonglong(0)
        lib().hs_render_bdpt(self.h, C.byref(cam), C.byref(params), int(pruned), film.ctypes.data_as(C.c_void_p),
                             C.byref(rays))
        return film, rays.value

    def debug_shading(self, what, inputs, iparam=0, cam=None):
        IN = {0: 10, 1: 10, 2: 10, 3: 7, 4: 7, 5: 4, 6: 2, 7: 13, 8: 5}
        OUT = {0: 9, 1: 9, 2: 2, 3: 10, 4: 12, 5: 5, 6: 1, 7: 13, 8: 8}
        a = np.ascontiguousarray(inputs, np.float32).reshape(-1, IN[what])
        out = np.zeros((len(a), OUT[what]), np.float32)
        lib().hs_debug_shading(self.h, C.byref(cam) if cam is not None else None, int(what), int(iparam),
                               a.ctypes.data_as(C.c_void_p), C.c_size_t(len(a)), out.ctypes.data_as(C.c_void_p))
        return out

    def set_rng_tape(self, tape, stride=0):
        """RNG replay for the next renders of ANY HostSim (process-global, like the device's per-scene tape)."""
        if tape is None:
            self._tape = None
            lib().hs_set_rng_tape(None, 0)
            return
        self._tape = np.ascontiguousarray(tape, np.float32).ravel()
        lib().hs_set_rng_tape(self._tape.ctypes.data_as(C.c_void_p), C.c_uint(int(stride)))

    def __del__(self):
        try:
            lib().hs_scene_destroy(self.h)
        except Exception:
            pass
