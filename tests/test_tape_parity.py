"""Image parity WITHOUT Monte-Carlo noise: the integrators replay the reference's own random numbers.

The reference draws every number of a render from one MT19937 stream.  oracle/ref_harness.cpp (PT) and the interposer
oracle/ref_hooks.cpp (BDPT) record, for every sample / sub-path of the UNMODIFIED reference, the stretch of the stream
it starts at; with that tape installed (wrt_debug_set_rng_tape / hostsim's set_rng_tape) our integrators draw the same
numbers in the same order, follow the same paths and must produce the same film — per pixel.

CPU tier (here): the host build of the device code (tests/hostsim) against (a) the committed golden tapes, always, and
(b) the live reference on more scenes when oracle/_ref is built.  IEEE float arithmetic without FMA on both sides and
the same libm, so the bar is the float-summation order of the film: 1e-6 relative.  The GPU tier is in
test_gpu_tape.py (CUDA's cosf / sinf / powf differ from libm by ulps, so a few paths may branch differently there)."""
import os

import numpy as np
import pytest

import scenes
import util
from hostsim_py import HostSim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(mine, film, rtol=2e-6):
    return np.abs(mine - film) <= rtol * (1.0 + np.abs(film))


def torus_small(res):
    sc, z = scenes.load_fixture("torus")
    sc.cam12 = sc.cam12.copy(); sc.cam12[9] = res; sc.cam12[10] = res; sc.width = sc.height = res
    return sc


def test_pt_golden_tape(wrt):
    z = np.load(os.path.join(GOLDEN, "tape_pt_small_mixed.npz"))
    res = int(z["res"])
    sc = scenes.small_mixed_scene(res, res)
    hs = util.host_scene(wrt, sc); sim = HostSim(hs.desc(), hs)
    sim.set_rng_tape(z["tape"], int(z["stride"]))
    try:
        mine, rays = sim.render_pt(hs.camera(), wrt.PtParams(res, res, int(z["spp"]), int(z["depth"]), 1, 0, 1, 0.0))
    finally:
        sim.set_rng_tape(None)
    assert z["film"].mean() > 0.1
    assert close(mine, z["film"]).all(), "max abs diff %g" % np.abs(mine - z["film"]).max()


def test_bdpt_golden_tape(wrt):
    z = np.load(os.path.join(GOLDEN, "tape_bdpt_small_mixed.npz"))
    res = int(z["res"])
    sc = scenes.small_mixed_scene(res, res)
    hs = util.host_scene(wrt, sc); sim = HostSim(hs.desc(), hs)
    sim.set_rng_tape(z["tape"], int(z["stride"]))
    try:
        mine, rays = sim.render_bdpt(hs.camera(), wrt.BdptParams(res, res, int(z["iterations"]), 0, 10, 3, 1, 0, 1, 1.0, 0))
    finally:
        sim.set_rng_tape(None)
    assert z["film"].mean() > 0.005
    assert close(mine, z["film"]).all(), "max abs diff %g" % np.abs(mine - z["film"]).max()


@pytest.mark.parametrize("name,spp,depth", [("cornell", 4, 5), ("small_mixed", 4, 5), ("torus", 1, 7), ("torus", 16, 7)])
def test_pt_follows_the_reference_path_for_path(wrt, have_ref, name, spp, depth):
    """PathIntegrator (surfaceIntegrator.cpp:14-46, pathIntegrator.cpp:29-148) incl. torus.scene = BASELINE config 0's
    scene (glass + diffuse + a far-away emitter): same random numbers -> same film, per pixel."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    res = 40
    sc = {"cornell": lambda: scenes.cornell_box_scene(res, res), "small_mixed": lambda: scenes.small_mixed_scene(res, res),
          "torus": lambda: torus_small(res)}[name]()
    hs = util.host_scene(wrt, sc); sim = HostSim(hs.desc(), hs)
    ref = util.ref_scene(sc)
    film, tape, rgb, draws = ref.render_pt_tape(spp, depth, seed=5489 + spp, stride=96)
    sim.set_rng_tape(tape, 96)
    try:
        mine, rays = sim.render_pt(hs.camera(), wrt.PtParams(res, res, spp, depth, 1, 0, 1, 0.0))
    finally:
        sim.set_rng_tape(None)
    assert film.mean() > 0 and draws.max() >= 11
    bad = ~close(mine, film).all(axis=2)
    assert bad.sum() == 0, "%d of %d pixels differ, max abs diff %g" % (bad.sum(), bad.size, np.abs(mine - film).max())


@pytest.mark.parametrize("name", ["cornell", "small_mixed"])
def test_bdpt_follows_the_reference_path_for_path(wrt, have_ref, name):
    """BidirPathTracing::runIteration (bidirPathTracing.cpp:53-265) with the shipped controlLength = 3 gating and every
    quirk of SURVEY App. C: light sub-paths, light tracing splats, camera sub-paths, direct illumination, connections."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    res, iters = 32, 2
    sc = scenes.cornell_box_scene(res, res) if name == "cornell" else scenes.small_mixed_scene(res, res)
    hs = util.host_scene(wrt, sc); sim = HostSim(hs.desc(), hs)
    ref = util.ref_scene(sc, "bdpt")
    film, tape, draws = ref.render_bdpt_tape(iters, seed=5489, stride=160)
    sim.set_rng_tape(tape, 160)
    try:
        mine, rays = sim.render_bdpt(hs.camera(), wrt.BdptParams(res, res, iters, 0, 10, 3, 1, 0, 1, 1.0, 0))
    finally:
        sim.set_rng_tape(None)
    assert film.mean() > 0.005
    bad = ~close(mine, film).all(axis=2)
    assert bad.sum() == 0, "%d of %d pixels differ, max abs diff %g" % (bad.sum(), bad.size, np.abs(mine - film).max())
