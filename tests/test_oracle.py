"""Pins the oracle: the plain-C port (oracle/port.c) against the golden vectors produced by the
reference itself, and — when the compiled reference (oracle/_ref) is present — against it directly
on adversarial batches.  Bit-exact bar (integer ids, flags, float bit patterns)."""
import numpy as np
import pytest

import engines
import scenes
import util

FIXTURES = ["torus", "cbox_dragon", "bunny", "small_mixed", "mixed_torus"]


@pytest.mark.parametrize("name", FIXTURES)
def test_port_matches_golden(wrt, name):
    sc, z = scenes.load_fixture(name)
    engines.check_against_golden(wrt, engines.PortEngine(wrt, sc), sc, z)


@pytest.mark.parametrize("name", FIXTURES)
def test_reference_matches_golden(wrt, have_ref, name):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    sc, z = scenes.load_fixture(name)
    engines.check_against_golden(wrt, engines.RefEngine(wrt, sc), sc, z)


@pytest.mark.parametrize("name", ["torus", "small_mixed"])
def test_port_matches_reference_adversarial(wrt, have_ref, name):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    sc = scenes.load_fixture(name)[0] if name != "small_mixed" else scenes.small_mixed_scene()
    port = engines.PortEngine(wrt, sc); ref = engines.RefEngine(wrt, sc)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 12000))
    a = port.intersect(rays, full=True); b = ref.intersect(rays, full=True)
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(util.bits(a[1]), util.bits(b[1]))
    hit = a[0] >= 0
    assert np.array_equal(util.bits(a[2][hit]), util.bits(b[2][hit]))      # p
    assert np.array_equal(util.bits(a[3][hit]), util.bits(b[3][hit]))      # n
    assert np.array_equal(a[4][hit], b[4][hit]) and np.array_equal(a[5][hit], b[5][hit])


def test_primitive_kats(have_ref):
    """T0: Triangle::hit / Sphere::hit / AABB::hit known-answer tests, port vs compiled reference."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    import ctypes as C
    from oracle import refpy, portpy
    R, P = refpy.lib(), portpy.lib()
    rng = np.random.Generator(np.random.PCG64(11))
    f32p = C.POINTER(C.c_float)
    n = 4000
    for i in range(n):
        tri = rng.normal(size=9).astype(np.float32)
        o = (rng.normal(size=3) * 3).astype(np.float32)
        mode = i % 5
        if mode == 0:   d = (tri[:3] - o)                                   # through a vertex
        elif mode == 1: d = ((tri[:3] + tri[3:6]) * 0.5 - o)                # through an edge
        elif mode == 2: d = np.array([0.0, -0.0, 1.0]) * (1 if i % 2 else -1)  # axis parallel, signed zero
        elif mode == 3: d = (tri[3:6] - tri[:3])                            # parallel to an edge (edge-on)
        else:           d = rng.normal(size=3)
        d = d.astype(np.float32); d = d / np.float32(np.sqrt((d * d).sum(dtype=np.float32)))
        ray = np.concatenate([o, d, [0.0, 1e7]]).astype(np.float32)
        t1, t2 = C.c_float(), C.c_float(); p3 = (C.c_float * 3)(); n3 = (C.c_float * 3)(); ins = C.c_int()
        h1 = R.ref_triangle_hit(tri.ctypes.data_as(f32p), ray.ctypes.data_as(f32p), C.byref(t1), p3, n3, C.byref(ins))
        h2 = P.port_triangle_hit(tri.ctypes.data_as(f32p), ray.ctypes.data_as(f32p), C.byref(t2))
        assert h1 == h2 and np.float32(t1.value).view(np.uint32) == np.float32(t2.value).view(np.uint32)
        cr = np.concatenate([rng.normal(size=3), [abs(rng.normal()) + (1e-4 if i % 7 == 0 else 0.05)]]).astype(np.float32)
        i1, i2 = C.c_int(), C.c_int()
        h1 = R.ref_sphere_hit(cr.ctypes.data_as(f32p), ray.ctypes.data_as(f32p), C.byref(t1), p3, n3, C.byref(i1))
        h2 = P.port_sphere_hit(cr.ctypes.data_as(f32p), ray.ctypes.data_as(f32p), C.byref(t2), C.byref(i2))
        assert h1 == h2 and np.float32(t1.value).view(np.uint32) == np.float32(t2.value).view(np.uint32)
        if h1: assert i1.value == i2.value
        box = np.sort(rng.normal(size=(2, 3)).astype(np.float32), axis=0).ravel()
        a1, a2, b1, b2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        h1 = R.ref_aabb_hit(box.ctypes.data_as(f32p), ray.ctypes.data_as(f32p), C.byref(a1), C.byref(a2))
        h2 = P.port_aabb_hit(box.ctypes.data_as(f32p), ray.ctypes.data_as(f32p), C.byref(b1), C.byref(b2))
        assert h1 == h2
        if h1: assert a1.value == b1.value and a2.value == b2.value


@pytest.mark.parametrize("name", ["torus", "small_mixed"])
def test_port_shadow_any_and_grazing_match_reference(wrt, have_ref, name):
    """Pins the round-2 additions of the port: Scene::shadowRayTest with off-surface targets, bool Scene::intersect, and the
    grazing-ray class (rays in triangle planes, where Triangle::hit divides rounding noise by rounding noise): bit equality."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    sc = scenes.load_fixture(name)[0] if name != "small_mixed" else scenes.small_mixed_scene()
    port = engines.PortEngine(wrt, sc); ref = engines.RefEngine(wrt, sc)
    rays = wrt.make_rays(np.concatenate([engines.adversarial_rays(sc, 8000, seed=21), engines.grazing_rays(sc, 12000, seed=22, top_fraction=1.0)]))
    a, b = port.intersect(rays), ref.intersect(rays)
    assert np.array_equal(a[0], b[0]) and np.array_equal(util.bits(a[1]), util.bits(b[1]))
    tgt = engines.shadow_test_queries(port, rays)
    vis = port.shadow_test(rays, tgt)
    assert np.array_equal(vis, ref.shadow_test(rays, tgt)) and 0.2 < vis.mean() < 0.9
    assert np.array_equal(port.intersect_any(rays), ref.intersect_any(rays))
