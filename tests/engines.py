"""Uniform 'engine' adapters so one parity routine can check the oracle port, the compiled
reference, the hostsim build and the CUDA path (through the C ABI) against the golden vectors."""
import numpy as np

import scenes
import util


class PortEngine:
    name = "oracle-port"
    def __init__(self, wrt, sc):
        self.port, self.hs = util.port_scene(wrt, sc)
    def intersect(self, rays, full=False):
        return self.port.intersect(rays, full=full)
    def occluded(self, q9):
        return self.port.occluded(q9)


class RefEngine:
    name = "reference"
    def __init__(self, wrt, sc):
        self.ref = util.ref_scene(sc)
    def intersect(self, rays, full=False):
        return self.ref.intersect(rays, full=full)
    def occluded(self, q9):
        return self.ref.occluded(q9)


class HostSimEngine:
    def __init__(self, wrt, sc, pruned):
        from hostsim_py import HostSim
        self.name = "hostsim-" + ("pruned" if pruned else "exact")
        self.hs = util.host_scene(wrt, sc)
        self.sim = HostSim(self.hs.desc(), self.hs)
        self.pruned = pruned
    def intersect(self, rays, full=False):
        return self.sim.trace_closest_full(rays, self.pruned) if full else self.sim.trace_closest(rays, self.pruned)
    def occluded(self, q9):
        return self.sim.trace_occluded(q9, self.pruned)


class CudaEngine:
    def __init__(self, wrt, sc, pruned):
        self.name = "cuda-" + ("pruned" if pruned else "exact")
        self.hs = util.host_scene(wrt, sc)
        self.scene = wrt.Scene(self.hs)
        self.scene.set_traversal(wrt.TRAVERSE_PRUNED if pruned else wrt.TRAVERSE_EXACT)
    def intersect(self, rays, full=False):
        return self.scene.intersect(rays, full=full)
    def occluded(self, q9):
        return self.scene.occluded(q9)


def check_against_golden(wrt, engine, sc, z):
    """Batches P (primary), S (NEE occlusion) and R (secondary) of tests/golden/make_golden.py.
    Bar: primitive ids and occlusion flags identical; t bit-identical (tolerance 0 ulp)."""
    cam = wrt.Camera.from_ref_array(z["cam45"])
    rays = wrt.generate_rays(cam, scenes.pixel_centres(512, 512, step=2))
    prim, t, p, n, ins, mat = engine.intersect(rays, full=True)
    bad = np.nonzero(prim != z["P_prim"])[0]
    assert len(bad) == 0, "%s: %d primary prim-id mismatches, first at ray %d" % (engine.name, len(bad), bad[0])
    assert np.array_equal(util.bits(t), util.bits(z["P_t"])), "%s: primary t not bit-identical" % engine.name
    q = scenes.nee_queries(p, (prim >= 0) & (mat > 0), sc.lights)
    assert len(q) == int(z["S_n"])
    occ = engine.occluded(q)
    assert np.array_equal(np.packbits(occ), z["S_occ"]), "%s: occlusion flags differ" % engine.name
    r2 = wrt.make_rays(scenes.bounce_rays(p, n, prim >= 0))
    prim2, t2 = engine.intersect(r2)
    assert np.array_equal(prim2, z["R_prim"]), "%s: secondary prim ids differ" % engine.name
    assert np.array_equal(util.bits(t2), util.bits(z["R_t"])), "%s: secondary t not bit-identical" % engine.name


def adversarial_rays(sc, n=20000, seed=3):
    """Edge cases: axis-parallel directions (d = +-0 components), origins exactly on vertex coordinates
    (split planes are vertex coordinates), rays along shared edges, tiny t, rays from inside the box."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tri = sc.data[sc.kind == 0]
    lo = sc.data[sc.kind == 0].reshape(-1, 3).min(0); hi = sc.data[sc.kind == 0].reshape(-1, 3).max(0)
    ext = hi - lo
    out = []
    # (a) axis-parallel rays through random vertices
    v = tri.reshape(-1, 3)[rng.integers(0, len(tri) * 3, n // 4)]
    for k in range(len(v)):
        ax = k % 3; sign = 1.0 if (k // 3) % 2 == 0 else -1.0
        d = np.zeros(3, np.float32); d[ax] = sign
        if (k // 6) % 2: d[(ax + 1) % 3] = -0.0
        o = v[k].copy(); o[ax] = lo[ax] - sign * 0.25 * ext[ax] if sign > 0 else hi[ax] + 0.25 * ext[ax]
        out.append(np.concatenate([o, d]))
    # (b) rays aimed exactly at vertices and edge mid-points from outside
    m = n // 4
    t_idx = rng.integers(0, len(tri), m)
    tgt = np.where((np.arange(m) % 2 == 0)[:, None], tri[t_idx, 0:3], (tri[t_idx, 0:3] + tri[t_idx, 3:6]) * np.float32(0.5))
    o = (lo + ext * (rng.random((m, 3)) * 3 - 1)).astype(np.float32)
    out.extend(np.concatenate([o, (tgt - o).astype(np.float32)], 1))
    # (c) origins ON triangles (t ~ 0 self hits) with random directions
    m = n // 4
    t_idx = rng.integers(0, len(tri), m)
    b = rng.random((m, 2)).astype(np.float32); b[b.sum(1) > 1] = 1 - b[b.sum(1) > 1]
    o = (tri[t_idx, 0:3] + (tri[t_idx, 3:6] - tri[t_idx, 0:3]) * b[:, :1] + (tri[t_idx, 6:9] - tri[t_idx, 0:3]) * b[:, 1:]).astype(np.float32)
    d = rng.normal(size=(m, 3)).astype(np.float32)
    out.extend(np.concatenate([o, d], 1))
    # (d) random rays from inside the scene box
    m = n - len(out)
    o = (lo + ext * rng.random((m, 3))).astype(np.float32)
    d = rng.normal(size=(m, 3)).astype(np.float32)
    out.extend(np.concatenate([o, d], 1))
    return np.asarray(out, np.float32)
