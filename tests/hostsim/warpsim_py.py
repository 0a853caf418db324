"""ctypes front-end of the TEST-ONLY warpsim library (tests/hostsim/warpsim.cpp): the product's warp-level traversal
schedulers executed on the CPU, one thread per lane."""
import ctypes as C
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_lib = None


def lib():
    global _lib
    if _lib is None:
        spec = importlib.util.spec_from_file_location("wrt_build", os.path.join(_ROOT, "winmad-s-raytracer-v1.0_b200", "build.py"))
        b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
        _lib = C.CDLL(b.build_warpsim())
    return _lib


class WarpSim:
    """sched: 3 = pooled (the default scheduler of the kernels), 2 = lane refill + vote."""

    def __init__(self, desc, keep=None):
        self._keep = keep
        self.h = C.c_void_p()
        err = C.create_string_buffer(256)
        if lib().ws_scene_create(C.byref(desc), C.byref(self.h), err):
            raise RuntimeError(err.value.decode())

    def trace_closest(self, rays8, pruned=True, sched=3):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        prim = np.full(len(r), -7, np.int32); t = np.zeros(len(r), np.float32)
        lib().ws_trace_closest(self.h, r.ctypes.data_as(C.c_void_p), C.c_size_t(len(r)), int(bool(pruned)), int(sched),
                               prim.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
        return prim, t

    @staticmethod
    def par_stats():
        """(finished by the whole-warp traversal, handed back to the ordinary rounds) since the last call."""
        out = (C.c_ulonglong * 2)()
        lib().ws_par_stats(out)
        return int(out[0]), int(out[1])

    def trace_occluded(self, q9, pruned=True, sched=3):
        q = np.ascontiguousarray(q9, np.float32).reshape(-1, 9)
        occ = np.full(len(q), 7, np.uint8)
        lib().ws_trace_occluded(self.h, q.ctypes.data_as(C.c_void_p), C.c_size_t(len(q)), int(bool(pruned)), int(sched),
                                occ.ctypes.data_as(C.c_void_p))
        return occ

    def __del__(self):
        try:
            lib().ws_scene_destroy(self.h)
        except Exception:
            pass
