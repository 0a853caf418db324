"""Shared helpers for the tests (scene plumbing between fixtures, oracle, hostsim and the C ABI)."""
import hashlib
import os

import numpy as np

import scenes


def tree_digest(t):
    """Same digest tests/golden/make_golden.py stores (topology, split planes, leaf lists)."""
    h = hashlib.sha256()
    inner = t["axis"] >= 0
    for a in (t["axis"], t["split"][inner], t["left"], t["right"], t["first_ref"][~inner], t["nref"], t["refs"]):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def host_scene(wrt, sc):
    return wrt.HostScene.from_arrays(sc.materials, sc.kind, sc.data, sc.matid, sc.lights, sc.cam12, build=True)


def ref_scene(sc, kind="pt"):
    from oracle import refpy
    r = refpy.RefScene(kind)
    r.build(sc.materials, sc.kind, sc.data, sc.matid, sc.lights, sc.cam12, sc.width, sc.height)
    return r


def port_scene(wrt, sc, hs=None):
    """Oracle port over OUR host-built tree (the tree itself is checked against the reference's)."""
    from oracle import portpy
    hs = hs or host_scene(wrt, sc)
    a = hs.arrays()
    return portpy.PortScene(a["prim_kind"], a["prim_data"], a["prim_matid"], a["tree"]), hs


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def golden_batches(wrt, sc, z):
    """Re-create the ray batches of make_golden.py from the fixture (camera matrices from the reference)."""
    from oracle import portpy
    cam = wrt.Camera.from_ref_array(z["cam45"])
    xy = scenes.pixel_centres(512, 512, step=2)
    rays_p = wrt.generate_rays(cam, xy)
    return cam, rays_p


def rel_rmse(a, b):
    """Per-pixel RMSE relative to the mean radiance of the reference image b (north_star bound: 1 %)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-12))
