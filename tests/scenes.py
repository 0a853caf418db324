"""Test / bench scenes as plain arrays (no file formats, no reference paths at run time).

Bundled scenes (configs C1/C2 of SURVEY.md §8d) are read from the committed fixtures under
tests/golden/ (generated from the reference by tests/golden/make_golden.py).  Synthetic scenes
(configs C3-C5) are generated procedurally here.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# material rows: diffuse rgb, phong rgb, phongExp, specular rgb, index
def material(diffuse=(0, 0, 0), phong=(0, 0, 0), phong_exp=1.0, specular=(0, 0, 0), index=-1.0):
    return list(diffuse) + list(phong) + [phong_exp] + list(specular) + [index]


class SceneArrays:
    def __init__(self, name, kind, data, matid, materials, lights, cam12, width, height):
        self.name = name
        self.kind = np.ascontiguousarray(kind, np.int32)
        self.data = np.ascontiguousarray(data, np.float32).reshape(-1, 9)
        self.matid = np.ascontiguousarray(matid, np.int32)
        self.materials = np.ascontiguousarray(materials, np.float32).reshape(-1, 11)
        self.lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 12)
        self.cam12 = np.ascontiguousarray(cam12, np.float32)
        self.width, self.height = int(width), int(height)

    @property
    def n_prims(self):
        return len(self.kind)


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = SceneArrays(name, z["kind"], z["data"], z["matid"], z["materials"], z["lights"], z["cam12"],
                     int(z["width"]), int(z["height"]))
    return sc, z


def with_light_quad(tris, matid, light_quad, intensity):
    """Append a 2-triangle area light (quad corners a,b,c,d) — matId -(f+1) as scene.cpp:427."""
    a, b, c, d = [np.asarray(v, np.float32) for v in light_quad]
    lt = np.stack([np.concatenate([a, b, c]), np.concatenate([a, c, d])]).astype(np.float32)
    lights = np.concatenate([lt, np.tile(np.asarray(intensity, np.float32), (2, 1))], axis=1)
    data = np.concatenate([tris, lt]).astype(np.float32)
    mat = np.concatenate([matid, np.array([-1, -2], np.int32)]).astype(np.int32)
    return data, mat, lights


def displaced_torus(n, R=1.0):
    """n x n tessellated torus with r = .35 + .05 sin 7u cos 5v: 2 n^2 triangles."""
    u = (np.arange(n, dtype=np.float64) / n) * 2 * np.pi
    v = (np.arange(n, dtype=np.float64) / n) * 2 * np.pi
    U, V = np.meshgrid(u, v, indexing="ij")
    r = 0.35 + 0.05 * np.sin(7 * U) * np.cos(5 * V)
    P = np.stack([(R + r * np.cos(V)) * np.cos(U), (R + r * np.cos(V)) * np.sin(U), r * np.sin(V)], -1).astype(np.float32)
    i0 = np.arange(n); i1 = (i0 + 1) % n
    a = P[i0][:, i0]; b = P[i1][:, i0]; c = P[i1][:, i1]; d = P[i0][:, i1]
    t1 = np.concatenate([a, b, c], -1).reshape(-1, 9)
    t2 = np.concatenate([a, c, d], -1).reshape(-1, 9)
    return np.concatenate([t1, t2]).astype(np.float32)


def synthetic_torus_scene(n=708, width=1920, height=1080, n_spheres=0, seed=1, name=None, floor=False, sphere_radius=None):
    """C3 (n=708 -> 1 002 528 triangles) / C5 (n=2237 + 100k spheres) of SURVEY.md §8d: the displaced
    torus, one diffuse material, a 2-triangle area light above, optional random spheres.  `floor=True`
    adds a 2-triangle ground quad; the reference's SAH builder degenerates into x-slabs with such huge
    primitives (thousands of leaves per ray), so it is off for the benchmark configs."""
    tris = displaced_torus(n)
    floor_z = -0.45
    f = 3.0
    data = tris
    matid = np.full(len(tris), 1, np.int32)
    if floor:
        fl = np.array([[-f, -f, floor_z, f, -f, floor_z, f, f, floor_z],
                       [-f, -f, floor_z, f, f, floor_z, -f, f, floor_z]], np.float32)
        data = np.concatenate([tris, fl])
        matid = np.concatenate([matid, np.full(2, 2, np.int32)])
    kind = np.zeros(len(data), np.int32)
    if n_spheres:
        rng = np.random.Generator(np.random.PCG64(seed))
        lo, hi = tris.reshape(-1, 3).min(0).astype(np.float64), tris.reshape(-1, 3).max(0).astype(np.float64)
        c = lo + (hi - lo) * rng.random((n_spheres, 3))
        ext = tris.reshape(-1, 3).max(0) - tris.reshape(-1, 3).min(0)
        scene_r = 0.5 * float(np.linalg.norm(ext))             # radius of the mesh's bounding sphere (~2.0)
        rad = scene_r * 10 ** (-3 + rng.random(n_spheres))     # log-uniform in [1e-3, 1e-2] * scene radius
        if sphere_radius is not None:                           # absolute radii (the default ones are below the reference's
            rad = sphere_radius[0] + (sphere_radius[1] - sphere_radius[0]) * rng.random(n_spheres)   # hittable size, sphere.cpp:44-46)
        sph = np.zeros((n_spheres, 9), np.float32)
        sph[:, :3] = c; sph[:, 3] = rad
        data = np.concatenate([data, sph]); kind = np.concatenate([kind, np.ones(n_spheres, np.int32)])
        matid = np.concatenate([matid, np.full(n_spheres, 3, np.int32)])
    h = 2.2
    # wound so that the emitting side (normal d1 x d2) faces down
    quad = [(-0.6, -0.6, h), (-0.6, 0.6, h), (0.6, 0.6, h), (0.6, -0.6, h)]
    lt_data, lt_mat, lights = with_light_quad(np.zeros((0, 9), np.float32), np.zeros(0, np.int32), quad, (40, 40, 40))
    data = np.concatenate([data, lt_data]); matid = np.concatenate([matid, lt_mat])
    kind = np.concatenate([kind, np.zeros(2, np.int32)])
    materials = [material(), material(diffuse=(0.75, 0.55, 0.35)), material(diffuse=(0.7, 0.7, 0.7)),
                 material(diffuse=(0.2, 0.3, 0.8), phong=(0.3, 0.3, 0.3), phong_exp=20.0)]
    pos = np.array([2.6, -2.6, 1.7], np.float32)
    fwd = -pos / np.linalg.norm(pos)
    up = np.array([0, 0, 1], np.float32)
    # resolution: <height> -> xResolution, <width> -> yResolution in the loader (scene.cpp:292-295);
    # PT maps raster x to the column index, so xResolution is the image width here.
    cam12 = np.concatenate([pos, fwd, up, [width, height, 40.0]]).astype(np.float32)
    return SceneArrays(name or ("synthetic_torus_%d" % n), kind, data, matid, materials, lights, cam12, width, height)


def cornell_box_scene(width=512, height=512, closed=True, name="cornell"):
    """C4: closed Cornell-box-style scene with two boxes and a 2-triangle area light (BDPT)."""
    def quad(a, b, c, d):
        a, b, c, d = [np.asarray(v, np.float32) for v in (a, b, c, d)]
        return [np.concatenate([a, b, c]), np.concatenate([a, c, d])]
    L = 1.0
    tris, mats = [], []
    def add(q, m):
        tris.extend(q); mats.extend([m, m])
    add(quad((-L, -L, -L), (L, -L, -L), (L, L, -L), (-L, L, -L)), 1)       # floor  (z=-L)
    add(quad((-L, -L, L), (-L, L, L), (L, L, L), (L, -L, L)), 1)           # ceiling
    add(quad((-L, L, -L), (L, L, -L), (L, L, L), (-L, L, L)), 1)           # back wall (y=+L)
    add(quad((-L, -L, -L), (-L, L, -L), (-L, L, L), (-L, -L, L)), 2)       # left (green)
    add(quad((L, -L, -L), (L, -L, L), (L, L, L), (L, L, -L)), 3)           # right (red)
    if closed:
        add(quad((-L, -L, -L), (-L, -L, L), (L, -L, L), (L, -L, -L)), 1)   # front wall behind the camera
    def box(c, hx, hy, hz, m):
        cx, cy, cz = c
        x0, x1, y0, y1, z0, z1 = cx - hx, cx + hx, cy - hy, cy + hy, cz - hz, cz + hz
        add(quad((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)), m)
        add(quad((x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0)), m)
        add(quad((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)), m)
        add(quad((x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0)), m)
        add(quad((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)), m)
        add(quad((x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1)), m)
    box((-0.35, 0.3, -0.4), 0.3, 0.3, 0.6, 1)
    box((0.4, -0.25, -0.7), 0.3, 0.3, 0.3, 4)
    data = np.stack(tris).astype(np.float32)
    matid = np.array(mats, np.int32)
    z = L - 0.01
    lq = [(-0.25, -0.25, z), (-0.25, 0.25, z), (0.25, 0.25, z), (0.25, -0.25, z)]
    data, matid, lights = with_light_quad(data, matid, lq, (25, 25, 25))
    kind = np.zeros(len(data), np.int32)
    materials = [material(), material(diffuse=(0.75, 0.75, 0.75)), material(diffuse=(0.16, 0.80, 0.17)),
                 material(diffuse=(0.80, 0.15, 0.15)), material(diffuse=(0.1, 0.1, 0.1), phong=(0.7, 0.7, 0.7), phong_exp=90.0)]
    cam12 = np.array([0, -0.98, 0, 0, 1, 0, 0, 0, 1, width, height, 90.0], np.float32)
    return SceneArrays(name, kind, data, matid, materials, lights, cam12, width, height)


def mixed_torus_scene(width=512, height=512):
    """3 200-triangle displaced torus + 200 spheres large enough to be hit (r in [0.05, 0.15]): a tree of > 512 nodes, so
    the GPU runs its pooled scheduler, with sphere records and skip records in the leaves.  Golden fixture `mixed_torus`."""
    return synthetic_torus_scene(n=40, width=width, height=height, n_spheres=200, sphere_radius=(0.05, 0.15), name="mixed_torus")


def small_mixed_scene(width=64, height=64):
    """Tiny scene with triangles AND spheres (incl. a glass and a mirror sphere) for smoke / edge tests."""
    sc = cornell_box_scene(width, height, closed=True, name="small_mixed")
    sph = np.zeros((3, 9), np.float32)
    sph[0, :4] = [0.0, 0.1, 0.45, 0.25]
    sph[1, :4] = [-0.55, -0.45, -0.75, 0.22]
    sph[2, :4] = [0.55, 0.5, 0.2, 0.18]
    data = np.concatenate([sc.data[:-2], sph, sc.data[-2:]])
    kind = np.concatenate([sc.kind[:-2], np.ones(3, np.int32), sc.kind[-2:]])
    matid = np.concatenate([sc.matid[:-2], np.array([5, 6, 1], np.int32), sc.matid[-2:]])
    materials = np.concatenate([sc.materials, np.array([material(specular=(1, 1, 1), index=1.5),
                                                         material(specular=(1, 1, 1))], np.float32)])
    return SceneArrays("small_mixed", kind, data, matid, materials, sc.lights, sc.cam12, width, height)


# ---- deterministic ray batches (numpy float32 +,-,*,/,sqrt only: bit-reproducible) -------------------
def pixel_centres(width, height, step=1):
    ii, jj = np.meshgrid(np.arange(0, height, step), np.arange(0, width, step), indexing="ij")
    return np.stack([jj.ravel(), ii.ravel()], 1).astype(np.float32)   # raster (x=j, y=i), A.6


def _hash_u32(x):
    x = np.asarray(x, np.uint32).copy()
    x ^= x >> np.uint32(16); x *= np.uint32(0x7feb352d); x ^= x >> np.uint32(15)
    x *= np.uint32(0x846ca68b); x ^= x >> np.uint32(16)
    return x


def unit_floats(n, salt):
    """n floats in [0,1) with 24-bit mantissas, from an integer hash."""
    with np.errstate(over="ignore"):
        h = _hash_u32(np.arange(n, dtype=np.uint32) * np.uint32(0x9E3779B9) + np.uint32(salt))
    return (h & np.uint32(0xffffff)).astype(np.float32) / np.float32(16777216.0)


def nee_queries(p, n_hits_mask, lights12, ks=((.25, .25), (.5, .5), (.75, .25), (.1, .9))):
    """Batch S: occluded(hit + d*EPS, d, hit + d*(dist-EPS)) toward sampleTriangle(u_k) on light 0
    (pathIntegrator.cpp:95-96), all in float32."""
    f = np.float32
    eps = f(1e-3)
    l = lights12[0].astype(np.float32)
    p0, p1, p2 = l[0:3], l[3:6], l[6:9]
    out = []
    hp = p[n_hits_mask].astype(np.float32)
    for (ux, uy) in ks:
        u1 = np.sqrt(f(ux)); beta = f(1) - u1; gamma = f(uy) * u1
        lp = (p0 + (p1 - p0) * beta + (p2 - p0) * gamma).astype(np.float32)
        d = (lp[None, :] - hp).astype(np.float32)
        dist = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(np.float32)).astype(np.float32)
        d = (d / dist[:, None]).astype(np.float32)
        p1q = (hp + d * eps).astype(np.float32)
        p2q = (hp + d * (dist - eps)[:, None]).astype(np.float32)
        out.append(np.concatenate([p1q, d, p2q], 1))
    return np.concatenate(out).astype(np.float32)


def bounce_rays(p, n, mask, salt=7):
    """Batch R: secondary rays leaving the hit points into the hemisphere of n (no trig: dir = n + s,
    |s| < 1, normalised by the Ray constructor later).  Returns (m,6) origin+dir."""
    f = np.float32
    hp = p[mask].astype(np.float32); hn = n[mask].astype(np.float32)
    m = len(hp)
    s = np.stack([unit_floats(m, salt), unit_floats(m, salt + 1), unit_floats(m, salt + 2)], 1)
    s = ((s * f(2) - f(1)) * f(0.55)).astype(np.float32)
    d = (hn + s).astype(np.float32)
    o = (hp + d * f(1e-3)).astype(np.float32)
    return np.concatenate([o, d], 1).astype(np.float32)


# ---- degenerate inputs ------------------------------------------------------------------------------------------
def edge_scenes():
    """Degenerate inputs the reference accepts silently: a single primitive, zero-area and duplicated triangles, a sliver,
    coordinates of very different magnitude, a sphere inside a triangle soup."""
    base = cornell_box_scene(32, 32)
    light = base.data[base.matid < 0]
    def mk(name, tris, extra_kind=None, extra=None):
        tris = np.asarray(tris, np.float32).reshape(-1, 9)
        data = tris; kind = np.zeros(len(tris), np.int32); matid = np.ones(len(tris), np.int32)
        if extra is not None:
            data = np.concatenate([data, extra]); kind = np.concatenate([kind, extra_kind]); matid = np.concatenate([matid, np.ones(len(extra), np.int32)])
        data = np.concatenate([data, light]); kind = np.concatenate([kind, np.zeros(len(light), np.int32)])
        matid = np.concatenate([matid, base.matid[base.matid < 0]])
        return SceneArrays(name, kind, data.astype(np.float32), matid.astype(np.int32), base.materials, base.lights, base.cam12, 32, 32)
    one = [[-0.5, 0.2, -0.5, 0.5, 0.2, -0.5, 0.0, 0.2, 0.6]]
    yield mk("single", one)
    yield mk("zero_area", one + [[0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1], [0, 0, 0, 1, 1, 1, 2, 2, 2]])
    yield mk("duplicates", one * 40)
    yield mk("sliver", one + [[-0.9, 0.5, 0.0, 0.9, 0.5, 0.0, 0.9, 0.5, 1e-6]])
    rng = np.random.Generator(np.random.PCG64(9))
    soup = (rng.random((300, 9)) * 1.6 - 0.8).astype(np.float32)
    yield mk("soup_and_spheres", soup, np.ones(3, np.int32), np.array([[0, 0.3, 0, 0.3, 0, 0, 0, 0, 0], [0.4, 0.1, -0.3, 0.2, 0, 0, 0, 0, 0],
                                                                         [-0.3, -0.2, 0.4, 0.05, 0, 0, 0, 0, 0]], np.float32))
    big = soup.copy(); big[:150] *= np.float32(500.0)
    yield mk("mixed_magnitudes", big)
    # enough primitives for a tree of > 512 nodes (the GPU's pooled scheduler): small random triangles, a few huge ones
    c = (rng.random((4000, 1, 3)) * 2 - 1).astype(np.float32)
    many = (c + (rng.random((4000, 3, 3)).astype(np.float32) - 0.5) * np.float32(0.08)).reshape(-1, 9)
    many[:6] = (rng.random((6, 9)) * 4 - 2).astype(np.float32)
    yield mk("many_small_few_huge", many)



# ---- .scene + OBJ files in the reference's format (for the loader / CLI tests) -----------------------------
def write_scene_files(sc, out_dir, name="scene"):
    """Writes `sc` (triangles only) as <out_dir>/<name>.scene plus one OBJ per material and one for the light
    (format of R/torus.scene; loaded by Scene::loadScene and by our host/scene_io.cpp).  Returns the scene path."""
    import os
    assert (sc.kind == 0).all(), "the .scene format of the reference has no sphere element"
    def obj(path, tris):
        with open(path, "w") as f:
            for t in tris.reshape(-1, 3):
                f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in t))
            for k in range(len(tris)):
                f.write("f %d %d %d\n" % (3 * k + 1, 3 * k + 2, 3 * k + 3))
    xml = ["<scene>", "<camera>",
           '<position x="%.9g" y="%.9g" z="%.9g"/>' % tuple(float(x) for x in sc.cam12[0:3]),
           '<forward x="%.9g" y="%.9g" z="%.9g"/>' % tuple(float(x) for x in sc.cam12[3:6]),
           '<up x="%.9g" y="%.9g" z="%.9g"/>' % tuple(float(x) for x in sc.cam12[6:9]),
           '<resolution height="%d" width="%d"/>' % (int(sc.cam12[9]), int(sc.cam12[10])),   # the loader swaps the two (scene.cpp:292-295)
           '<horizontalFOV horizontalFOV="%.9g"/>' % float(sc.cam12[11]), "</camera>"]
    for m in np.asarray(sc.materials, np.float32).reshape(-1, 11):
        xml.append('<material><diffuse r="%.9g" g="%.9g" b="%.9g"/><glossy r="%.9g" g="%.9g" b="%.9g"/>'
                   '<specular r="%.9g" g="%.9g" b="%.9g"/><phongExp phongExp="%.9g"/><refracIndex refracIndex="%.9g"/></material>'
                   % (m[0], m[1], m[2], m[3], m[4], m[5], m[7], m[8], m[9], m[6], m[10]))
    for mid in sorted(set(int(x) for x in sc.matid if x > 0)):
        path = os.path.join(out_dir, "%s_mat%d.obj" % (name, mid))
        obj(path, sc.data[sc.matid == mid])
        xml.append('<object><file_path path="%s"/><matid matid="%d"/></object>' % (path, mid))
    lpath = os.path.join(out_dir, "%s_light.obj" % name)
    obj(lpath, sc.data[sc.matid < 0])
    inten = np.asarray(sc.lights, np.float32).reshape(-1, 12)[0, 9:12]
    xml.append('<area_light><file_path path="%s"/><intensity r="%.9g" g="%.9g" b="%.9g"/></area_light>' % (lpath, inten[0], inten[1], inten[2]))
    xml.append("</scene>")
    spath = os.path.join(out_dir, name + ".scene")
    open(spath, "w").write("\n".join(xml) + "\n")
    return spath
