"""The CUDA kernels' per-ray code (csrc/traverse.cuh) compiled for the CPU ("hostsim", test-only) must
match the golden vectors bit for bit in both traversal modes, and EXACT must equal PRUNED — the same
checks the GPU tier runs on the real kernels, runnable where there is no GPU."""
import numpy as np
import pytest

import engines
import scenes
import util


@pytest.mark.parametrize("pruned", [False, True])
@pytest.mark.parametrize("name", ["torus", "cbox_dragon", "bunny", "small_mixed", "mixed_torus"])
def test_hostsim_matches_golden(wrt, name, pruned):
    sc, z = scenes.load_fixture(name)
    engines.check_against_golden(wrt, engines.HostSimEngine(wrt, sc, pruned), sc, z)


@pytest.mark.parametrize("name", ["torus", "small_mixed", "synthetic"])
def test_hostsim_exact_equals_pruned_and_port(wrt, name):
    if name == "torus": sc = scenes.load_fixture(name)[0]
    elif name == "small_mixed": sc = scenes.small_mixed_scene()
    else: sc = scenes.synthetic_torus_scene(n=64, width=64, height=64, n_spheres=500)
    ex = engines.HostSimEngine(wrt, sc, False); pr = engines.HostSimEngine(wrt, sc, True)
    port = engines.PortEngine(wrt, sc)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 40000))
    a, b, c = ex.intersect(rays, full=True), pr.intersect(rays, full=True), port.intersect(rays, full=True)
    for x in (b, c):
        assert np.array_equal(a[0], x[0])
        assert np.array_equal(util.bits(a[1]), util.bits(x[1]))
        hit = a[0] >= 0
        assert np.array_equal(util.bits(a[2][hit]), util.bits(x[2][hit]))
        assert np.array_equal(util.bits(a[3][hit]), util.bits(x[3][hit]))
        assert np.array_equal(a[4][hit], x[4][hit]) and np.array_equal(a[5][hit], x[5][hit])


@pytest.mark.parametrize("knobs", [("2", "2"), ("5", "3"), ("16", "6")])
def test_hostsim_skip_record_layouts(wrt, monkeypatch, knobs):
    """Leaf skip records (scene_layout.cpp) at aggressive settings — every leaf of >= 2 entries chunked, nested
    groups everywhere — must not change a single result: EXACT (ignores them) == PRUNED (jumps over prunable
    chunks) == oracle port, on adversarial rays with infinite AND finite ray.tmax (the reference ends the whole
    traversal at the first popped node with ray.tmax < tmin, KDtreeAccel.cpp:323)."""
    monkeypatch.setenv("WRT_LEAF_SKIP_MIN", knobs[0]); monkeypatch.setenv("WRT_LEAF_SKIP_CHUNK", knobs[1])
    sc = scenes.synthetic_torus_scene(n=64, width=64, height=64, n_spheres=500)
    ex = engines.HostSimEngine(wrt, sc, False); pr = engines.HostSimEngine(wrt, sc, True)
    monkeypatch.setenv("WRT_LEAF_SKIP", "0")
    plain = engines.HostSimEngine(wrt, sc, True)
    assert pr.sim.num_recs() > plain.sim.num_recs()          # the layout really contains skip records
    port = engines.PortEngine(wrt, sc)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 30000, seed=7))
    short = rays.copy(); short[:, 7] = np.random.default_rng(7).uniform(0.05, 4.0, len(rays)).astype(np.float32)
    for rr in (rays, short):
        want = port.intersect(rr)
        for e in (ex, pr, plain):
            got = e.intersect(rr)
            assert np.array_equal(got[0], want[0]) and np.array_equal(util.bits(got[1]), util.bits(want[1])), e.name
    q = scenes.nee_queries(*(lambda a: (a[2], (a[0] >= 0) & (a[5] > 0)))(pr.intersect(rays, full=True)), sc.lights)
    assert np.array_equal(pr.occluded(q), port.occluded(q)) and np.array_equal(ex.occluded(q), port.occluded(q))


def test_nan_interval_terminates_like_the_reference(wrt):
    """Regression (found by the adversarial batch while testing a child-pair scheduler): an axis-parallel ray whose
    origin lies exactly on split planes produces NaN kd intervals; the reference then pops an entry with
    tmin = +inf, sees ray.tmax < tmin and BREAKS, leaving a closer primitive in a later entry untested."""
    sc = scenes.synthetic_torus_scene(n=96, width=64, height=64, n_spheres=2000)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 60000))[13280:13300]
    port = engines.PortEngine(wrt, sc)
    want = port.intersect(rays)
    assert want[0][9] == 965                                   # the reference's (non-closest) answer for ray 13289
    for pruned in (False, True):
        got = engines.HostSimEngine(wrt, sc, pruned).intersect(rays)
        assert np.array_equal(got[0], want[0]) and np.array_equal(util.bits(got[1]), util.bits(want[1]))


def test_visit_counters_match_port(wrt):
    """Reference-semantics work counters (the B_ray inputs of the roofline) agree with the instrumented port."""
    sc, z = scenes.load_fixture("cbox_dragon")
    sim = engines.HostSimEngine(wrt, sc, False); port = engines.PortEngine(wrt, sc)
    cam = wrt.Camera.from_ref_array(z["cam45"])
    rays = wrt.generate_rays(cam, scenes.pixel_centres(512, 512, step=4))
    c1 = sim.sim.count_visits(rays, pruned=False)
    c2 = port.port.intersect(rays, count=True)[-1]
    assert c1 == c2
    c3 = sim.sim.count_visits(rays, pruned=True)
    assert c3["tri"] <= c1["tri"] and c3["leaf"] <= c1["leaf"]


def test_empty_and_ragged_inputs(wrt):
    sc = scenes.small_mixed_scene()
    e = engines.HostSimEngine(wrt, sc, True)
    prim, t = e.intersect(np.zeros((0, 8), np.float32))
    assert len(prim) == 0
    # rays that miss the root box entirely, rays with zero-length direction (NaN dir after the ctor)
    od = np.array([[10, 10, 10, 1, 0, 0], [0, 0, 0, 0, 0, 0], [0, -0.5, 0, 0, 1, 0]], np.float32)
    with np.errstate(all="ignore"):
        rays = wrt.make_rays(od)
    prim, t = e.intersect(rays)
    port = engines.PortEngine(wrt, sc)
    p2, t2 = port.intersect(rays)
    assert np.array_equal(prim, p2) and np.array_equal(util.bits(t), util.bits(t2))
    assert prim[0] == -1 and prim[2] >= 0


def test_hostsim_whitted_matches_reference(wrt, have_ref):
    """the per-node Whitted code (whitted_logic.cuh, compiled for the CPU) against the reference's
    WhittedIntegrator.  The reference image has NaN pixels by construction (0 * 0 / 0 for a light seen from its
    back side, whitted.cpp:37-38): the NaN masks must agree as well as two reference seeds agree with each other,
    and the finite pixels statistically."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from hostsim_py import HostSim
    res, spp = 48, 16
    for sc in (scenes.small_mixed_scene(res, res), scenes.cornell_box_scene(res, res)):
        hs = util.host_scene(wrt, sc); sim = HostSim(hs.desc(), hs)
        mine, rays = sim.render_whitted(hs.camera(), wrt.PtParams(res, res, spp, 7, 11, 0, 1, 0.0))
        ref = util.ref_scene(sc, "whitted"); ref.reset_traverse_calls()
        r1 = ref.render_whitted(spp, 7, seed=5489); calls = ref.traverse_calls(); r2 = ref.render_whitted(spp, 7, seed=31)
        n1, n2, nm = np.isnan(r1).any(2), np.isnan(r2).any(2), np.isnan(mine).any(2)
        assert n1.sum() > 0 and (n1 != nm).sum() <= 1.5 * (n1 != n2).sum() + 8
        ok = ~(n1 | n2 | nm)
        rm = (r1 + r2) * 0.5
        assert abs(mine[ok].mean() - rm[ok].mean()) <= 0.01 * rm[ok].mean()
        assert np.abs(mine[ok] - rm[ok]).mean() <= 1.1 * np.abs(r1[ok] - r2[ok]).mean()
        assert 0.7 * calls <= rays <= 1.001 * calls      # black connections are not traced; never more rays than the reference


def _block_mean(img, b):
    h, w, c = img.shape
    return img[: h // b * b, : w // b * b].reshape(h // b, b, w // b, b, c).mean(axis=(1, 3))


@pytest.mark.parametrize("kind", ["pt", "bdpt"])
def test_hostsim_integrators_match_reference_statistically(wrt, have_ref, kind):
    """T3 on the CPU: the kernels' per-path logic (hostsim) against the reference renderer on a small
    closed Cornell box.  Monte-Carlo estimators with different RNGs: the mean radiance must agree to
    1.5 % and the 8x8-block rRMSE must be at the reference's own two-seed noise floor."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from hostsim_py import HostSim
    res = 40
    sc = scenes.cornell_box_scene(res, res)
    hs = util.host_scene(wrt, sc); sim = HostSim(hs.desc(), hs); cam = hs.camera()
    if kind == "pt":
        spp = 64
        mine, rays = sim.render_pt(cam, wrt.PtParams(res, res, spp, 5, 11, 0, 1, 0.0))
        ref = util.ref_scene(sc, "pt")
        ref.reset_traverse_calls()
        r1 = ref.render_pt(spp, 5, seed=5489); calls = ref.traverse_calls(); r2 = ref.render_pt(spp, 5, seed=31)
        n_samples = res * res * spp
    else:
        iters = 48
        mine, rays = sim.render_bdpt(cam, wrt.BdptParams(res, res, iters, 0, 10, 3, 11, 0, 1, 0.0, 0))
        ref = util.ref_scene(sc, "bdpt")
        ref.reset_traverse_calls()
        r1 = ref.render_bdpt(iters, seed=5489) / iters; calls = ref.traverse_calls()
        r2 = ref.render_bdpt(iters, seed=31) / iters
        n_samples = res * res * iters
    rm = (r1 + r2) * 0.5
    floor = util.rel_rmse(_block_mean(r1, 8), _block_mean(r2, 8))
    err = util.rel_rmse(_block_mean(mine, 8), _block_mean(rm, 8))
    print("%s: mean %.5f vs %.5f, rRMSE %.4f floor %.4f, rays/sample %.2f vs %.2f"
          % (kind, mine.mean(), rm.mean(), err, floor, rays / n_samples, calls / n_samples))
    assert abs(mine.mean() - rm.mean()) <= 0.015 * rm.mean()
    assert err <= 1.3 * floor + 0.002
    # PT queues only the shadow rays that can contribute, BDPT skips the BSDF-sampled DI ray when the light
    # sample is occluded: never more rays than the reference, and not fewer than 70 % of them
    assert 0.7 * calls <= rays <= 1.001 * calls


def test_edge_scenes_tree_and_traversal(wrt, have_ref):
    """Builder and traversal on degenerate inputs: our KD-tree equals the reference's (live, where oracle/_ref is built),
    and EXACT == PRUNED == oracle port on adversarial rays."""
    for sc in scenes.edge_scenes():
        hs = util.host_scene(wrt, sc)
        a = hs.arrays()["tree"]
        if have_ref:
            b = util.ref_scene(sc).tree()
            for k in ("axis", "left", "right", "nref", "refs"):
                assert np.array_equal(a[k], b[k]), (sc.name, k)
            inner = b["axis"] >= 0
            assert np.array_equal(util.bits(a["split"][inner]), util.bits(b["split"][inner])), sc.name
        port = engines.PortEngine(wrt, sc)
        with np.errstate(all="ignore"):
            rays = wrt.make_rays(engines.adversarial_rays(sc, 4000, seed=2))
        want = port.intersect(rays)
        for pruned in (False, True):
            got = engines.HostSimEngine(wrt, sc, pruned).intersect(rays)
            assert np.array_equal(got[0], want[0]) and np.array_equal(util.bits(got[1]), util.bits(want[1])), (sc.name, pruned)


# ---- the warp-level schedulers on the CPU (tests/hostsim/warpsim.cpp) ------------------------------------------------
@pytest.mark.parametrize("sched", [3, 2])
def test_warpsim_schedulers_match_port(wrt, monkeypatch, sched):
    """The product's warp-level schedulers themselves — csrc/trace_pooled.cuh (scheduler 3: rings, refill, re-queueing, the
    global-scratch stacks, early exit of boolean queries) and csrc/trace_persistent.cuh (scheduler 2: lane refill + vote) —
    compiled for the CPU on an emulation of the warp primitives (one thread per lane, 4 warps) and run on the mixed
    triangle + sphere golden scene with every leaf chunked into skip records: closest hits (infinite and finite ray.tmax,
    PRUNED and EXACT) and occlusion flags must equal the oracle port's, and the reference's golden vectors."""
    from warpsim_py import WarpSim
    monkeypatch.setenv("WRT_LEAF_SKIP_MIN", "3"); monkeypatch.setenv("WRT_LEAF_SKIP_CHUNK", "2")
    sc, z = scenes.load_fixture("mixed_torus")
    hs = util.host_scene(wrt, sc)
    ws = WarpSim(hs.desc(), hs)
    port = engines.PortEngine(wrt, sc)
    n = 2500 if sched == 3 else 1200
    rays = wrt.make_rays(engines.adversarial_rays(sc, n, seed=4))
    short = rays.copy(); short[:, 7] = np.random.default_rng(4).uniform(0.05, 4.0, len(rays)).astype(np.float32)
    for rr in (rays, short):
        want = port.intersect(rr)
        for pruned in (True, False):
            got = ws.trace_closest(rr, pruned, sched)
            assert np.array_equal(got[0], want[0]) and np.array_equal(util.bits(got[1]), util.bits(want[1])), (sched, pruned)
    full = port.intersect(rays, full=True)
    q = scenes.nee_queries(full[2], (full[0] >= 0) & (full[5] > 0), sc.lights)[:n]
    assert np.array_equal(ws.trace_occluded(q, True, sched), port.occluded(q))
    if sched == 3:
        # small launches (fewer rays than pool slots: even-share refill, tail loop entered with up to 16 rays, per-ray lane groups
        # of 2 ... 32 lanes in the leaves): closest hits and occlusion flags
        occ = port.occluded(q)
        want_r = port.intersect(rays)
        ws.par_stats()
        for m in (3, 7, 40, 100, 230):
            for pruned in (True, False):
                got = ws.trace_closest(rays[:m], pruned, 3)
                assert np.array_equal(got[0], want_r[0][:m]) and np.array_equal(util.bits(got[1]), util.bits(want_r[1][:m])), (m, pruned)
            assert np.array_equal(ws.trace_occluded(q[:m], True, 3), occ[:m]), m
        done, back = ws.par_stats()
        print("whole-warp traversal of last rays: %d finished, %d handed back" % (done, back))
        assert done > 20            # the path under test was taken (par_traverse), and its answers are the oracle's
    # golden primary rays of the fixture (a strided sample): the reference's own answers
    cam = wrt.Camera.from_ref_array(z["cam45"])
    prim_rays = wrt.generate_rays(cam, scenes.pixel_centres(512, 512, step=2))
    sel = np.arange(0, len(prim_rays), 53)[: n]
    got = ws.trace_closest(prim_rays[sel], True, sched)
    assert np.array_equal(got[0], z["P_prim"][sel]) and np.array_equal(util.bits(got[1]), util.bits(z["P_t"][sel]))


def test_warpsim_whole_warp_traversal_of_last_rays(wrt, monkeypatch):
    """par_traverse (trace_pooled.cuh): the last rays of a launch are traversed by the whole warp — sub-trees side by side, far children
    handed to idle lanes, no acceptance rule, closest hit + runner-up; near-ties go back to the ordinary rounds.  Launches of one or
    two rays put every ray through it: adversarial rays (vertex / edge hits = near-ties, axis-parallel = degenerate, spheres) on the
    mixed golden scene with chunked leaves, PRUNED and EXACT, closest hits and occlusion flags — all must equal the oracle's."""
    from warpsim_py import WarpSim
    monkeypatch.setenv("WRT_LEAF_SKIP_MIN", "3"); monkeypatch.setenv("WRT_LEAF_SKIP_CHUNK", "2")
    sc, z = scenes.load_fixture("mixed_torus")
    hs = util.host_scene(wrt, sc)
    ws = WarpSim(hs.desc(), hs)
    port = engines.PortEngine(wrt, sc)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 260, seed=12))
    short = rays.copy(); short[:, 7] = np.random.default_rng(2).uniform(0.05, 4.0, len(rays)).astype(np.float32)
    ws.par_stats()
    for rr in (rays, short):
        want = port.intersect(rr)
        for pruned in (True, False):
            for i in range(0, len(rr), 2):
                got = ws.trace_closest(rr[i:i + 2], pruned, 3)
                assert np.array_equal(got[0], want[0][i:i + 2]) and np.array_equal(util.bits(got[1]), util.bits(want[1][i:i + 2])), (i, pruned)
    done, back = ws.par_stats()
    print("closest: %d rays finished by the whole warp, %d handed back (near-ties / the reference's tmax stop)" % (done, back))
    assert done > 350 and back > 10
    full = port.intersect(rays, full=True)
    q = scenes.nee_queries(full[2], (full[0] >= 0) & (full[5] > 0), sc.lights)[:160]
    occ = port.occluded(q)
    for i in range(0, len(q), 2):
        assert np.array_equal(ws.trace_occluded(q[i:i + 2], True, 3), occ[i:i + 2]), i
    done, back = ws.par_stats()
    print("occlusion: %d finished by the whole warp, %d handed back" % (done, back))
    assert done > 60


def test_warpsim_whole_warp_traversal_on_edge_scenes(wrt):
    """par_traverse on the degenerate inputs of scenes.edge_scenes() — duplicated triangles (two primitives with the SAME t: the
    runner-up test must send them back to the ordinary rounds, where the first one listed wins), zero-area triangles, a sliver,
    mixed magnitudes, a sphere in a triangle soup: launches of two rays, closest hits == oracle."""
    from warpsim_py import WarpSim
    ws_done = ws_back = 0
    for sc in scenes.edge_scenes():
        hs = util.host_scene(wrt, sc)
        ws = WarpSim(hs.desc(), hs)
        port = engines.PortEngine(wrt, sc)
        with np.errstate(all="ignore"):
            rays = wrt.make_rays(engines.adversarial_rays(sc, 80, seed=5))
        want = port.intersect(rays)
        ws.par_stats()
        for i in range(0, len(rays), 2):
            got = ws.trace_closest(rays[i:i + 2], True, 3)
            assert np.array_equal(got[0], want[0][i:i + 2]) and np.array_equal(util.bits(got[1]), util.bits(want[1][i:i + 2])), (sc.name, i)
        d, b = ws.par_stats(); ws_done += d; ws_back += b
    print("edge scenes: %d rays finished by the whole warp, %d handed back" % (ws_done, ws_back))
    assert ws_done > 100 and ws_back > 0


def test_warpsim_nan_interval_and_tiny_batches(wrt):
    """Pooled scheduler on the CPU: the NaN-interval regression rays (tests above) and batches smaller than a warp / than the
    refill threshold (1, 5, 33 rays) — exhaustion and partially filled rings."""
    from warpsim_py import WarpSim
    sc = scenes.synthetic_torus_scene(n=96, width=64, height=64, n_spheres=2000)
    hs = util.host_scene(wrt, sc)
    ws = WarpSim(hs.desc(), hs)
    port = engines.PortEngine(wrt, sc)
    rays = wrt.make_rays(engines.adversarial_rays(sc, 60000))[13270:13310]
    want = port.intersect(rays)
    assert want[0][19] == 965
    for m in (1, 5, 33, len(rays)):
        got = ws.trace_closest(rays[:m], True, 3)
        assert np.array_equal(got[0], want[0][:m]) and np.array_equal(util.bits(got[1]), util.bits(want[1][:m])), m
    assert len(ws.trace_closest(rays[:0], True, 3)[0]) == 0


@pytest.mark.parametrize("name", ["synthetic", "torus", "synthetic_far"])
def test_grazing_rays_on_the_host_build(wrt, name):
    """CPU twin of test_gpu_traversal.py::test_grazing_rays_exact_is_exact_and_pruned_differs_only_in_plane: EXACT == oracle on rays
    in triangle planes; every PRUNED disagreement is a near-plane event (|cos| <= 2e-4 to a winner's normal)."""
    import engines
    base = scenes.synthetic_torus_scene(n=48, width=64, height=64, n_spheres=300) if name.startswith("synthetic") else scenes.load_fixture("torus")[0]
    sc = engines.far_scene(base) if name.endswith("_far") else base
    rays = wrt.make_rays(engines.grazing_rays(sc, 60000, seed=31))
    port = engines.PortEngine(wrt, sc)
    want = port.intersect(rays)
    ex = engines.HostSimEngine(wrt, sc, False).intersect(rays)
    assert np.array_equal(ex[0], want[0]) and np.array_equal(util.bits(ex[1]), util.bits(want[1]))
    pr = engines.HostSimEngine(wrt, sc, True).intersect(rays)
    bad = np.nonzero(pr[0] != want[0])[0]
    assert 0 < len(bad) <= 0.005 * len(rays)          # the generator does reach the class (if it stops doing so, it has lost its point)
    tri = sc.data.astype(np.float64)
    worst = 0.0
    for i in bad:
        cs = []
        for p in (want[0][i], pr[0][i]):
            if p >= 0 and sc.kind[p] == 0:
                T = tri[p]; n = np.cross(T[3:6] - T[0:3], T[6:9] - T[0:3]); n /= np.linalg.norm(n)
                cs.append(abs(float(n @ rays[i, 3:6].astype(np.float64))))
        worst = max(worst, min(cs))
    assert worst <= 2e-4, worst
