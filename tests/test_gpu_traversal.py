"""GPU tier: the CUDA kernels, called through the C ABI (include/wrt.h), against the golden vectors,
the oracle (compiled reference when it travelled with the repo, else the plain-C port) and themselves
(EXACT == PRUNED on large batches).  Bar: primitive ids / flags identical, t bit-identical."""
import numpy as np
import pytest

import engines
import scenes
import util

pytestmark = pytest.mark.gpu
FIXTURES = ["torus", "cbox_dragon", "bunny", "small_mixed", "mixed_torus"]


@pytest.mark.parametrize("pruned", [False, True])
@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_matches_golden(wrt, name, pruned):
    sc, z = scenes.load_fixture(name)
    engines.check_against_golden(wrt, engines.CudaEngine(wrt, sc, pruned), sc, z)


@pytest.mark.parametrize("name", FIXTURES)
def test_cuda_matches_oracle_full_batches(wrt, have_ref, name):
    """T1 at the full C2 batch size: 512^2 primary rays, 4 NEE queries per hit, secondary rays."""
    sc, z = scenes.load_fixture(name)
    oracle = engines.RefEngine(wrt, sc) if have_ref else engines.PortEngine(wrt, sc)
    cuda = engines.CudaEngine(wrt, sc, True)
    cam = wrt.Camera.from_ref_array(z["cam45"])
    rays = wrt.generate_rays(cam, scenes.pixel_centres(512, 512))
    a, b = cuda.intersect(rays, full=True), oracle.intersect(rays, full=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(util.bits(a[1]), util.bits(b[1]))
    hit = a[0] >= 0
    for k in (2, 3):
        assert np.array_equal(util.bits(a[k][hit]), util.bits(b[k][hit]))
    assert np.array_equal(a[4][hit], b[4][hit]) and np.array_equal(a[5][hit], b[5][hit])
    q = scenes.nee_queries(a[2], hit & (a[5] > 0), sc.lights)
    assert np.array_equal(cuda.occluded(q), oracle.occluded(q))
    r2 = wrt.make_rays(scenes.bounce_rays(a[2], a[3], hit))
    c, d = cuda.intersect(r2), oracle.intersect(r2)
    assert np.array_equal(c[0], d[0]) and np.array_equal(util.bits(c[1]), util.bits(d[1]))
    # the other query flavours of the seam
    assert np.array_equal(cuda.scene.intersect_any(rays), (a[0] >= 0).astype(np.uint8))
    vis = cuda.scene.shadowRayTest(rays[hit][:5000], a[2][hit][:5000])
    assert np.all(vis == 1.0)     # a ray's own hit point is "visible" (scene.cpp:64-67)


@pytest.mark.parametrize("name", ["torus", "small_mixed", "synthetic"])
def test_cuda_adversarial_vs_port(wrt, name):
    if name == "torus": sc = scenes.load_fixture(name)[0]
    elif name == "small_mixed": sc = scenes.small_mixed_scene()
    else: sc = scenes.synthetic_torus_scene(n=96, width=64, height=64, n_spheres=2000)
    port = engines.PortEngine(wrt, sc)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 60000))
    b = port.intersect(rays, full=True)
    for pruned in (False, True):
        a = engines.CudaEngine(wrt, sc, pruned).intersect(rays, full=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(util.bits(a[1]), util.bits(b[1]))
        hit = a[0] >= 0
        assert np.array_equal(util.bits(a[2][hit]), util.bits(b[2][hit]))
        assert np.array_equal(util.bits(a[3][hit]), util.bits(b[3][hit]))
        assert np.array_equal(a[4][hit], b[4][hit]) and np.array_equal(a[5][hit], b[5][hit])


@pytest.mark.parametrize("knobs", [("2", "2"), ("5", "3"), ("16", "6")])
def test_cuda_skip_record_layouts(wrt, monkeypatch, knobs):
    """Leaf skip records at aggressive settings (every leaf chunked, nested groups): CUDA EXACT == CUDA PRUNED ==
    oracle port on adversarial rays with infinite and finite ray.tmax, and on the occlusion query."""
    monkeypatch.setenv("WRT_LEAF_SKIP_MIN", knobs[0]); monkeypatch.setenv("WRT_LEAF_SKIP_CHUNK", knobs[1])
    sc = scenes.synthetic_torus_scene(n=96, width=64, height=64, n_spheres=2000)
    port = engines.PortEngine(wrt, sc)
    ex = engines.CudaEngine(wrt, sc, False); pr = engines.CudaEngine(wrt, sc, True)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 120000, seed=11))
    short = rays.copy(); short[:, 7] = np.random.default_rng(11).uniform(0.05, 4.0, len(rays)).astype(np.float32)
    for rr in (rays, short):
        want = port.intersect(rr)
        for e in (ex, pr):
            got = e.intersect(rr)
            assert np.array_equal(got[0], want[0]) and np.array_equal(util.bits(got[1]), util.bits(want[1])), e.name
    full = pr.intersect(rays, full=True)
    q = scenes.nee_queries(full[2], (full[0] >= 0) & (full[5] > 0), sc.lights)
    want = port.occluded(q)
    assert np.array_equal(pr.occluded(q), want) and np.array_equal(ex.occluded(q), want)


def test_cuda_nan_interval_terminates_like_the_reference(wrt):
    """See tests/test_hostsim.py::test_nan_interval_terminates_like_the_reference — same rays through the kernels,
    embedded in a batch large enough to run through the pooled scheduler."""
    sc = scenes.synthetic_torus_scene(n=96, width=64, height=64, n_spheres=2000)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 60000))
    port = engines.PortEngine(wrt, sc)
    want = port.intersect(rays)
    assert want[0][13289] == 965
    for pruned in (False, True):
        got = engines.CudaEngine(wrt, sc, pruned).intersect(rays)
        assert np.array_equal(got[0], want[0]) and np.array_equal(util.bits(got[1]), util.bits(want[1]))


@pytest.mark.parametrize("name", ["torus", "cbox_dragon", "bunny", "synthetic_1m"])
def test_exact_equals_pruned_large(wrt, name):
    """T1b: >= 1e7 rays per scene (primary at 4 sub-pixel offsets, secondary from the hits, random)."""
    if name == "synthetic_1m":
        sc = scenes.synthetic_torus_scene(n=708, width=1920, height=1080)
        cam = wrt.camera_setup(sc.cam12[0:3], sc.cam12[3:6], sc.cam12[6:9], sc.cam12[9], sc.cam12[10], sc.cam12[11])
        w, h = 1920, 1080
    else:
        sc, z = scenes.load_fixture(name)
        cam = wrt.Camera.from_ref_array(z["cam45"]); w, h = 512, 512
    hs = util.host_scene(wrt, sc)
    scene = wrt.Scene(hs)
    total = 0
    rng = np.random.Generator(np.random.PCG64(5))
    lo = sc.data[sc.kind == 0].reshape(-1, 3).min(0); hi = sc.data[sc.kind == 0].reshape(-1, 3).max(0)
    batches = []
    for off in ((0, 0), (.25, .25), (-.25, .4), (.4, -.3)):
        batches.append(("primary", wrt.generate_rays(cam, scenes.pixel_centres(w, h) + np.float32(off))))
    n_rand = 3_000_000
    o = (lo + (hi - lo) * (rng.random((n_rand, 3)) * 1.4 - 0.2)).astype(np.float32)
    d = rng.normal(size=(n_rand, 3)).astype(np.float32)
    batches.append(("random", wrt.make_rays(np.concatenate([o, d], 1))))
    for label, rays in batches:
        scene.set_traversal(wrt.TRAVERSE_EXACT); a = scene.intersect(rays, full=True)
        scene.set_traversal(wrt.TRAVERSE_PRUNED); b = scene.intersect(rays)
        assert np.array_equal(a[0], b[0]), (name, label, int((a[0] != b[0]).sum()))
        assert np.array_equal(util.bits(a[1]), util.bits(b[1])), (name, label)
        total += len(rays)
        if label == "primary":   # secondary rays from these hits, both modes
            hit = a[0] >= 0
            for salt in (7, 19, 31):
                r2 = wrt.make_rays(scenes.bounce_rays(a[2], a[3], hit, salt=salt))
                scene.set_traversal(wrt.TRAVERSE_EXACT); c = scene.intersect(r2)
                scene.set_traversal(wrt.TRAVERSE_PRUNED); e = scene.intersect(r2)
                assert np.array_equal(c[0], e[0]) and np.array_equal(util.bits(c[1]), util.bits(e[1])), (name, "secondary")
                total += len(r2)
    assert total >= (1e7 if w > 512 else 4e6)


def test_headline_scene_against_oracle_port(wrt):
    """C3's own scene (1 002 528 triangles, reference-exact KD build on the host) against the oracle port on a
    sample of primary, secondary and adversarial rays the port finishes in seconds."""
    sc = scenes.synthetic_torus_scene(n=708, width=1920, height=1080)
    port = engines.PortEngine(wrt, sc)
    scene = wrt.Scene(port.hs)
    cam = port.hs.camera()
    rays = wrt.generate_rays(cam, scenes.pixel_centres(1920, 1080, step=12) + np.float32(0.37))
    a = scene.intersect(rays, full=True); b = port.intersect(rays, full=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(util.bits(a[1]), util.bits(b[1]))
    hit = a[0] >= 0
    assert 0.2 < hit.mean() < 0.9
    for k in (2, 3):
        assert np.array_equal(util.bits(a[k][hit]), util.bits(b[k][hit]))
    r2 = wrt.make_rays(scenes.bounce_rays(a[2], a[3], hit))
    c, d = scene.intersect(r2), port.intersect(r2)
    assert np.array_equal(c[0], d[0]) and np.array_equal(util.bits(c[1]), util.bits(d[1]))
    q = scenes.nee_queries(a[2], hit, sc.lights)
    assert np.array_equal(scene.occluded(q), port.occluded(q))
    adv = wrt.make_rays(engines.adversarial_rays(sc, 6000))
    e, f = scene.intersect(adv), port.intersect(adv)
    assert np.array_equal(e[0], f[0]) and np.array_equal(util.bits(e[1]), util.bits(f[1]))


def test_c5_class_scene_against_oracle_port_and_exact(wrt):
    """C5's generator at 1/5 scale (2 000 000 triangles + 20 000 spheres: long leaves, skip records with groups, sphere
    records): CUDA PRUNED == oracle port on a sample the port finishes in seconds, and PRUNED == EXACT on 4K-frame
    primary + secondary batches."""
    sc = scenes.synthetic_torus_scene(n=1000, width=3840, height=2160, n_spheres=20000)
    port = engines.PortEngine(wrt, sc)
    scene = wrt.Scene(port.hs)
    cam = port.hs.camera()
    rays = wrt.generate_rays(cam, scenes.pixel_centres(3840, 2160, step=24) + np.float32(0.41))
    a = scene.intersect(rays, full=True); b = port.intersect(rays, full=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(util.bits(a[1]), util.bits(b[1]))
    hit = a[0] >= 0
    r2 = wrt.make_rays(scenes.bounce_rays(a[2], a[3], hit))
    c, d = scene.intersect(r2), port.intersect(r2)
    assert np.array_equal(c[0], d[0]) and np.array_equal(util.bits(c[1]), util.bits(d[1]))
    # rays aimed at sphere centres from the camera and from the opposite side.  Reference quirk: Sphere::hit rejects
    # t_hc = r^2 - d_perp^2 unless it exceeds EPS = 1e-3 (sphere.cpp:44-46), so spheres with r < 0.0316 — all of C5's —
    # can never be hit; they still cost their tests.  Both sides must agree on exactly that.
    cen = sc.data[sc.kind == 1][:4000, 0:3]
    o1 = np.broadcast_to(sc.cam12[0:3].astype(np.float32), cen.shape)
    o2 = (2 * cen.mean(0) - sc.cam12[0:3]).astype(np.float32)[None, :] + np.zeros_like(cen)
    aim = wrt.make_rays(np.concatenate([np.concatenate([o1, cen - o1], 1), np.concatenate([o2, cen - o2], 1)]).astype(np.float32))
    g, h = scene.intersect(aim), port.intersect(aim)
    assert np.array_equal(g[0], h[0]) and np.array_equal(util.bits(g[1]), util.bits(h[1]))
    assert (sc.kind[g[0][g[0] >= 0]] == 1).sum() == 0 and float(sc.data[sc.kind == 1][:, 3].max()) < 0.0316
    q = scenes.nee_queries(a[2], hit & (a[5] > 0), sc.lights)
    assert np.array_equal(scene.occluded(q), port.occluded(q))
    adv = wrt.make_rays(engines.adversarial_rays(sc, 6000))
    e, f = scene.intersect(adv), port.intersect(adv)
    assert np.array_equal(e[0], f[0]) and np.array_equal(util.bits(e[1]), util.bits(f[1]))
    # size-independent property at the full frame: PRUNED (pooled scheduler, skip records) == EXACT
    big = wrt.generate_rays(cam, scenes.pixel_centres(3840, 2160, step=2))
    p1 = scene.intersect(big, full=True)
    big2 = wrt.make_rays(scenes.bounce_rays(p1[2], p1[3], p1[0] >= 0))
    scene.set_traversal(wrt.TRAVERSE_EXACT)
    x1 = scene.intersect(big); x2 = scene.intersect(big2)
    scene.set_traversal(wrt.TRAVERSE_PRUNED)
    p2 = scene.intersect(big2)
    assert np.array_equal(p1[0], x1[0]) and np.array_equal(util.bits(p1[1]), util.bits(x1[1]))
    assert np.array_equal(p2[0], x2[0]) and np.array_equal(util.bits(p2[1]), util.bits(x2[1]))


def test_cuda_edge_scenes(wrt):
    """Degenerate inputs (single primitive, zero-area / duplicated triangles, sliver, mixed magnitudes, soup + spheres):
    CUDA EXACT == CUDA PRUNED == oracle port, closest hits and occlusion flags."""
    for sc in scenes.edge_scenes():
        port = engines.PortEngine(wrt, sc)
        with np.errstate(all="ignore"):
            rays = wrt.make_rays(engines.adversarial_rays(sc, 20000, seed=2))
        want = port.intersect(rays, full=True)
        q = scenes.nee_queries(want[2], (want[0] >= 0) & (want[5] > 0), sc.lights)
        occ = port.occluded(q)
        for pruned in (False, True):
            e = engines.CudaEngine(wrt, sc, pruned)
            got = e.intersect(rays)
            assert np.array_equal(got[0], want[0]) and np.array_equal(util.bits(got[1]), util.bits(want[1])), (sc.name, pruned)
            assert np.array_equal(e.occluded(q), occ), (sc.name, pruned)


def test_visit_counters_and_stats(wrt):
    sc, z = scenes.load_fixture("torus")
    cuda = engines.CudaEngine(wrt, sc, True); port = engines.PortEngine(wrt, sc)
    cam = wrt.Camera.from_ref_array(z["cam45"])
    rays = wrt.generate_rays(cam, scenes.pixel_centres(512, 512, step=4))
    c = cuda.scene.count_visits(rays)
    assert c == port.port.intersect(rays, count=True)[-1]
    cuda.scene.reset_stats()
    cuda.scene.intersect(rays)
    s = cuda.scene.stats()
    assert s.closest_rays == len(rays) and s.kernel_launches == 1 and s.last_trace_ms > 0


def test_empty_ragged_and_device_pointer_variants(wrt):
    import torch
    sc = scenes.small_mixed_scene()
    cuda = engines.CudaEngine(wrt, sc, True); port = engines.PortEngine(wrt, sc)
    prim, t = cuda.intersect(np.zeros((0, 8), np.float32))
    assert len(prim) == 0
    for n in (1, 31, 33, 1000):                      # ragged sizes around the warp width
        rays = wrt.make_rays(engines.adversarial_rays(sc, max(n, 8)))[:n]
        a, b = cuda.intersect(rays), port.intersect(rays)
        assert np.array_equal(a[0], b[0]) and np.array_equal(util.bits(a[1]), util.bits(b[1]))
    rays = wrt.make_rays(engines.adversarial_rays(sc, 5000))
    d_rays = torch.from_numpy(rays).cuda()
    d_prim = torch.empty(len(rays), dtype=torch.int32, device="cuda"); d_t = torch.empty(len(rays), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    cuda.scene.intersect_dev(d_rays.data_ptr(), len(rays), d_prim.data_ptr(), d_t.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    b = port.intersect(rays)
    assert np.array_equal(d_prim.cpu().numpy(), b[0]) and np.array_equal(util.bits(d_t.cpu().numpy()), util.bits(b[1]))


@pytest.mark.gpu
def test_small_launches_through_the_pooled_scheduler(wrt):
    """Launches of 1 ... 3 000 rays on a tree large enough for the pooled scheduler (torus.scene, 11 429 nodes): even-share refill,
    tail loop from the first round, lane groups in the leaves, whole-warp traversal of the last rays — closest hits (PRUNED and EXACT)
    and occlusion flags must equal the oracle's bit for bit.  (GPU twin of the warp-simulator tests in tests/test_hostsim.py.)"""
    sc, z = scenes.load_fixture("torus")
    port = engines.PortEngine(wrt, sc)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 3000, seed=21))
    want = port.intersect(rays)
    full = port.intersect(rays, full=True)
    q = scenes.nee_queries(full[2], (full[0] >= 0) & (full[5] > 0), sc.lights)
    occ = port.occluded(q)
    for pruned in (True, False):
        cuda = engines.CudaEngine(wrt, sc, pruned)
        for m in (1, 2, 7, 40, 230, 3000):
            got = cuda.intersect(rays[:m])
            assert np.array_equal(got[0], want[0][:m]) and np.array_equal(util.bits(got[1]), util.bits(want[1][:m])), (pruned, m)
            k = min(m, len(q))
            assert np.array_equal(cuda.occluded(q[:k]), occ[:k]), (pruned, m)


# ---- round 2: the two query flavours that were only self-tested, and the grazing-ray class -------------------------------
@pytest.mark.parametrize("name", ["torus", "small_mixed", "synthetic"])
def test_shadow_and_any_queries_vs_oracle(wrt, have_ref, name):
    """wrt_trace_shadow == Scene::shadowRayTest(ray, p) and wrt_trace_any == bool Scene::intersect(ray) (scene.cpp:45-69) with
    targets ON and OFF the surfaces: the ray's own hit point moved by 0 ... +-4 EPS along one axis (both sides of the component-wise
    tolerance of Vector3 ==), points before and beyond the hit, random targets for rays that miss."""
    if name == "torus": sc = scenes.load_fixture(name)[0]
    elif name == "small_mixed": sc = scenes.small_mixed_scene()
    else: sc = scenes.synthetic_torus_scene(n=96, width=64, height=64, n_spheres=2000)
    port = engines.PortEngine(wrt, sc)
    od = np.concatenate([engines.adversarial_rays(sc, 60000, seed=21), engines.grazing_rays(sc, 20000, seed=22, top_fraction=1.0)])
    rays = wrt.make_rays(od)
    tgt = engines.shadow_test_queries(port, rays)
    want_vis, want_any = port.shadow_test(rays, tgt), port.intersect_any(rays)
    assert 0.2 < want_vis.mean() < 0.9 and 0.05 < want_any.mean() < 1.0          # both outcomes are exercised
    if have_ref:
        ref = engines.RefEngine(wrt, sc)
        assert np.array_equal(ref.shadow_test(rays, tgt), want_vis) and np.array_equal(ref.intersect_any(rays), want_any)
    ex = engines.CudaEngine(wrt, sc, False)
    assert np.array_equal(ex.shadow_test(rays, tgt), want_vis) and np.array_equal(ex.intersect_any(rays), want_any)
    pr = engines.CudaEngine(wrt, sc, True)
    sel = slice(0, 60000)                                                        # PRUNED: the non-grazing part must be identical ...
    assert np.array_equal(pr.shadow_test(rays, tgt)[sel], want_vis[sel]) and np.array_equal(pr.intersect_any(rays)[sel], want_any[sel])
    assert (pr.intersect_any(rays)[60000:] != want_any[60000:]).mean() <= 0.01   # ... the in-plane class is characterised below


@pytest.mark.parametrize("name", ["cornell", "torus", "synthetic", "cornell_far", "synthetic_far"])
def test_grazing_rays_exact_is_exact_and_pruned_differs_only_in_plane(wrt, name):
    """VERDICT r1 item 1e.  Rays lying in (or within 1e-3 rad of) the plane of a triangle: Triangle::hit's denominator is then
    rounding noise and the reference reports hits with arbitrary t on triangles the ray does not geometrically reach.
      * EXACT traversal reproduces every one of them (ids and t bit-identical to the oracle) — it is the mode whose contract
        is unconditional;
      * PRUNED skips sub-trees on GEOMETRIC grounds, so it cannot see a noise hit on a triangle behind the real hit.  It must be
        identical on every ray whose winners (in either mode) are met at |cos(angle to the triangle normal)| > 2e-4 (0.011 deg),
        i.e. all disagreements are near-plane events (DESIGN.md §2 derives the bound; measured on a B200: 0.11 - 0.15 % of these
        adversarial rays, the largest |cos| among them 5e-5); the test prints their rate."""
    base = {"cornell": lambda: scenes.cornell_box_scene(64, 64), "torus": lambda: scenes.load_fixture("torus")[0],
            "synthetic": lambda: scenes.synthetic_torus_scene(n=96, width=64, height=64, n_spheres=2000)}[name.split("_")[0]]()
    sc = engines.far_scene(base) if name.endswith("_far") else base
    port = engines.PortEngine(wrt, sc)
    rays = wrt.make_rays(np.concatenate([engines.grazing_rays(sc, 150000, seed=31), engines.grazing_rays(sc, 50000, seed=32, top_fraction=1.0)]))
    want = port.intersect(rays)
    ex = engines.CudaEngine(wrt, sc, False).intersect(rays)
    assert np.array_equal(ex[0], want[0]) and np.array_equal(util.bits(ex[1]), util.bits(want[1]))
    pr = engines.CudaEngine(wrt, sc, True).intersect(rays)
    bad = np.nonzero((pr[0] != want[0]) | (util.bits(pr[1]) != util.bits(want[1])))[0]
    tri = sc.data.astype(np.float64)

    def cos_to(prims, r):
        out = np.ones(len(r))
        ok = (prims >= 0) & (sc.kind[np.maximum(prims, 0)] == 0)
        T = tri[prims[ok]]
        n = np.cross(T[:, 3:6] - T[:, 0:3], T[:, 6:9] - T[:, 0:3]); n /= np.linalg.norm(n, axis=1, keepdims=True) + 1e-300
        out[ok] = np.abs(np.sum(n * r[ok, 3:6].astype(np.float64), axis=1))
        return out
    c = np.minimum(cos_to(want[0][bad], rays[bad]), cos_to(pr[0][bad], rays[bad]))
    print("%s: %d of %d grazing rays differ under PRUNED (%.4f %%), largest |cos| among them %.2e"
          % (name, len(bad), len(rays), 100.0 * len(bad) / len(rays), c.max() if len(c) else 0.0))
    assert len(bad) <= 0.005 * len(rays)
    assert (c <= 2e-4).all(), "a PRUNED / EXACT disagreement at |cos| = %.3e is not a near-plane event" % c.max()
