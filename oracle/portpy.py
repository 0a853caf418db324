"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the plain-C oracle port (oracle/port.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_ref", "libwrt_port.so")
_lib = None
_f32p, _i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)


class PortSceneStruct(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("kind", _i32p), ("data9", _f32p), ("matid", _i32p),
                ("n_nodes", C.c_int32), ("axis", _i32p), ("split", _f32p), ("left", _i32p), ("right", _i32p),
                ("first_ref", _i32p), ("n_ref", _i32p), ("refs", _i32p), ("root_box", C.c_float * 6)]


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "port.c")
        if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(PORT_SO)
    return _lib


class PortScene:
    """kind/data/matid = Scene::objs; tree = dict(axis, split, left, right, first_ref, nref, refs, root_box)."""

    def __init__(self, kind, data, matid, tree):
        k = dict(kind=np.ascontiguousarray(kind, np.int32), data=np.ascontiguousarray(data, np.float32).reshape(-1, 9),
                 matid=np.ascontiguousarray(matid, np.int32))
        for key in ("axis", "left", "right", "first_ref", "nref", "refs"):
            k[key] = np.ascontiguousarray(tree[key], np.int32)
        k["split"] = np.ascontiguousarray(tree["split"], np.float32)
        self._keep = k
        s = PortSceneStruct()
        s.n_prims = len(k["kind"]); s.kind = k["kind"].ctypes.data_as(_i32p); s.data9 = k["data"].ctypes.data_as(_f32p)
        s.matid = k["matid"].ctypes.data_as(_i32p); s.n_nodes = len(k["axis"])
        s.axis = k["axis"].ctypes.data_as(_i32p); s.split = k["split"].ctypes.data_as(_f32p)
        s.left = k["left"].ctypes.data_as(_i32p); s.right = k["right"].ctypes.data_as(_i32p)
        s.first_ref = k["first_ref"].ctypes.data_as(_i32p); s.n_ref = k["nref"].ctypes.data_as(_i32p)
        s.refs = k["refs"].ctypes.data_as(_i32p)
        rb = tree["root_box"] if "root_box" in tree else tree["box"][0]
        s.root_box[:] = [float(v) for v in np.asarray(rb).ravel()[:6]]
        self.s = s

    def intersect(self, rays8, full=False, count=False):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8); n = len(r)
        prim = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        cnt = (C.c_ulonglong * 4)() if count else None
        if full:
            p = np.zeros((n, 3), np.float32); nn = np.zeros((n, 3), np.float32)
            ins = np.zeros(n, np.int32); mat = np.zeros(n, np.int32)
            lib().port_intersect(C.byref(self.s), r.ctypes.data_as(_f32p), C.c_longlong(n), prim.ctypes.data_as(_i32p),
                                 t.ctypes.data_as(_f32p), p.ctypes.data_as(_f32p), nn.ctypes.data_as(_f32p),
                                 ins.ctypes.data_as(_i32p), mat.ctypes.data_as(_i32p), cnt)
            res = (prim, t, p, nn, ins, mat)
        else:
            lib().port_intersect(C.byref(self.s), r.ctypes.data_as(_f32p), C.c_longlong(n), prim.ctypes.data_as(_i32p),
                                 t.ctypes.data_as(_f32p), None, None, None, None, cnt)
            res = (prim, t)
        if count:
            return res + (dict(inner=cnt[0], leaf=cnt[1], tri=cnt[2], sphere=cnt[3], rays=n),)
        return res

    def occluded(self, q9):
        q = np.ascontiguousarray(q9, np.float32).reshape(-1, 9)
        occ = np.zeros(len(q), np.uint8)
        lib().port_occluded(C.byref(self.s), q.ctypes.data_as(_f32p), C.c_longlong(len(q)),
                            occ.ctypes.data_as(C.POINTER(C.c_uint8)))
        return occ


    def shadow_test(self, rays8, target3):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        p = np.ascontiguousarray(target3, np.float32).reshape(-1, 3)
        vis = np.zeros(len(r), np.float32)
        lib().port_shadow_test(C.byref(self.s), r.ctypes.data_as(_f32p), p.ctypes.data_as(_f32p), C.c_longlong(len(r)), vis.ctypes.data_as(_f32p))
        return vis

    def intersect_any(self, rays8):
        r = np.ascontiguousarray(rays8, np.float32).reshape(-1, 8)
        hit = np.zeros(len(r), np.uint8)
        lib().port_intersect_any(C.byref(self.s), r.ctypes.data_as(_f32p), C.c_longlong(len(r)), hit.ctypes.data_as(C.POINTER(C.c_uint8)))
        return hit


def make_rays(od6):
    od = np.ascontiguousarray(od6, np.float32).reshape(-1, 6)
    out = np.zeros((len(od), 8), np.float32)
    lib().port_make_rays(od.ctypes.data_as(_f32p), C.c_longlong(len(od)), out.ctypes.data_as(_f32p))
    return out
