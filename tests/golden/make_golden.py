"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference and oracle/_ref/libwrt_ref.so):

    python tests/golden/make_golden.py

For each bundled scene and for one programmatic triangle + sphere scene it stores, in
tests/golden/<name>.npz:
  * the scene as arrays exactly as the reference's loader produced them (Scene::objs order),
    materials, AreaLight constructor arguments, Camera::setup arguments and the camera matrices;
  * a SHA-256 of the reference's flattened KD-tree (topology, split planes, leaf lists);
  * batch P: primary rays through pixel centres (every 2nd pixel of the 512^2 raster): reference
    Scene::intersect prim index + t;
  * batch S: four NEE occlusion queries per primary hit on a material (Scene::occluded flags);
  * batch R: deterministic secondary rays from the hit points: prim index + t;
  * reference-semantics visit counts are NOT stored (they come from the instrumented port).
Nothing here is read by the product; tests/ compare the oracle port, the hostsim build and the CUDA
path against these vectors.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refpy  # noqa: E402
import scenes  # noqa: E402

OBJ = refpy.REF_ROOT + "/ObjFiles"
HERE = os.path.dirname(os.path.abspath(__file__))


def mat_xml(m):
    return ('<material><diffuse r="%g" g="%g" b="%g"/><glossy r="%g" g="%g" b="%g"/>'
            '<specular r="%g" g="%g" b="%g"/><phongExp phongExp="%g"/><refracIndex refracIndex="%g"/></material>\n'
            % (m[0], m[1], m[2], m[3], m[4], m[5], m[7], m[8], m[9], m[6], m[10]))


def scene_xml(cam, res, fov, materials, objects, light_obj, intensity):
    s = "<scene>\n<camera>\n"
    s += '<position x="%r" y="%r" z="%r"/>\n<forward x="%r" y="%r" z="%r"/>\n<up x="%r" y="%r" z="%r"/>\n' % tuple(cam)
    s += '<resolution height="%d" width="%d"/>\n<horizontalFOV horizontalFOV="%r"/>\n</camera>\n' % (res, res, fov)
    for m in materials:
        s += mat_xml(m)
    for path, mid in objects:
        s += '<object><file_path path="%s/%s"/><matid matid="%d"/></object>\n' % (OBJ, path, mid)
    s += '<area_light><file_path path="%s"/><intensity r="%g" g="%g" b="%g"/></area_light>\n' % ((light_obj,) + tuple(intensity))
    return s + "</scene>\n"


def tree_digest(t):
    h = hashlib.sha256()
    inner = t["axis"] >= 0
    for a in (t["axis"], t["split"][inner], t["left"], t["right"], t["first_ref"][~inner], t["nref"], t["refs"]):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def bunny_light_obj(tmp):
    """A 2-triangle light above the bunny (the bunny OBJ has no emitter)."""
    p = os.path.join(tmp, "bunny_light.obj")
    open(p, "w").write("v -8 22 -8\nv 8 22 -8\nv 8 22 8\nv -8 22 8\nf 1 2 3\nf 1 3 4\n")
    return p


def make(name, scene_file=None, arrays=None):
    """scene_file: a .scene the reference's own loader reads; arrays: a tests/scenes.py SceneArrays built through the
    same constructor calls the loader makes (ref_build_scene) — the only way to get spheres into a Scene."""
    ref = refpy.RefScene("pt")
    if arrays is not None:
        n = ref.build(arrays.materials, arrays.kind, arrays.data, arrays.matid, arrays.lights, arrays.cam12, 512, 512)
    else:
        n = ref.load_file(scene_file, 512, 512)
    kind, data, matid = ref.prims()
    cam = ref.camera()
    lights22 = ref.lights()
    # AreaLight ctor args back from (p0, d1, d2): p1 = p0 + d1 is NOT exactly invertible in float, so
    # take the light triangles from the emitter primitives instead (same vertices the loader passed).
    em = np.where(matid < 0)[0]
    lights12 = np.concatenate([data[em], lights22[:, 9:12]], axis=1).astype(np.float32)
    assert len(em) == len(lights22)
    tr = ref.tree()
    out = dict(kind=kind, data=data, matid=matid, materials=ref.materials(), lights=lights12,
               # Camera::setup does not store its FOV argument (the loader sets the member separately), so for a
               # programmatic scene the arguments themselves are kept
               cam12=(cam[:12].copy() if arrays is None else np.asarray(arrays.cam12, np.float32).copy()), cam45=cam, width=512, height=512,
               tree_sha=tree_digest(tr), tree_nodes=len(tr["axis"]), tree_refs=len(tr["refs"]),
               tree_depth=tr["depth"], root_box=tr["box"][0], scene_sphere=ref.scene_sphere())
    xy = scenes.pixel_centres(512, 512, step=2)
    rays = ref.generate_rays(xy)
    prim, t, p, nrm, ins, mat = ref.intersect(rays, full=True)
    out.update(P_prim=prim, P_t=t)
    on_mat = (prim >= 0) & (mat > 0)
    q = scenes.nee_queries(p, on_mat, lights12)
    out.update(S_occ=np.packbits(ref.occluded(q)), S_n=len(q))
    od = scenes.bounce_rays(p, nrm, prim >= 0)
    r2 = refpy.make_rays(od)
    prim2, t2 = ref.intersect(r2)
    out.update(R_prim=prim2, R_t=t2)
    print("%-8s prims %6d nodes %6d refs %7d depth %2d | P hit %.3f | S n %6d occluded %.3f | R hit %.3f"
          % (name, n, len(tr["axis"]), len(tr["refs"]), tr["depth"], (prim >= 0).mean(), len(q),
             ref.occluded(q).mean(), (prim2 >= 0).mean()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    tmp = tempfile.mkdtemp(prefix="wrt_golden_")
    make("torus", refpy.fixed_torus_scene(tmp))
    M = scenes.material
    cbox_mats = [M(), M(diffuse=(0.8, 0.8, 0.8)), M(diffuse=(0.156863, 0.803922, 0.172549)),
                 M(diffuse=(0.803922, 0.152941, 0.152941)),
                 M(diffuse=(0.1, 0.1, 0.1), phong=(0.7, 0.7, 0.7), phong_exp=90.0)]
    cbox_cam = (-0.0439815, -4.12529, 0.222539, 0.00688625, 0.998505, -0.0542161, 3.73896e-4, 0.0542148, 0.998529)
    xml = scene_xml(cbox_cam, 512, 45.0, cbox_mats,
                    [("cbox_floor.obj", 1), ("cbox_ceiling.obj", 1), ("cbox_back.obj", 1), ("cbox_greenwall.obj", 2),
                     ("cbox_redwall.obj", 3), ("test_out_dragon.obj", 4)], OBJ + "/cbox_luminaire.obj", (25, 25, 25))
    p = os.path.join(tmp, "cbox_dragon.scene"); open(p, "w").write(xml)
    make("cbox_dragon", p)
    bunny_mats = [M(), M(diffuse=(0.7, 0.6, 0.5))]
    # bunny bbox is [-18,18]x[-17.8,17.8]x[-13.9,13.9]; camera at ~bbox-diagonal distance (58)
    cam = (0.0, 0.0, 58.0, 0.0, 0.0, -1.0, 0.0, 1.0, 0.0)
    xml = scene_xml(cam, 512, 40.0, bunny_mats, [("bunny.obj", 1)], bunny_light_obj(tmp), (60, 60, 60))
    p = os.path.join(tmp, "bunny.scene"); open(p, "w").write(xml)
    make("bunny", p)
    # triangles AND spheres (a glass and a mirror one among them, radii 0.18-0.25: hittable, unlike C5's): pins Sphere::hit
    make("small_mixed", arrays=scenes.small_mixed_scene(512, 512))
    make("mixed_torus", arrays=scenes.mixed_torus_scene(512, 512))
    # keep the generated scene files next to the fixtures' provenance (tiny, text)
    for f in ("cbox_dragon.scene", "bunny.scene"):
        txt = open(os.path.join(tmp, f)).read().replace(tmp, "$TMP")
        open(os.path.join(HERE, f + ".txt"), "w").write(txt)


if __name__ == "__main__":
    main()
