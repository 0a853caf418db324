"""Deterministic known-answer tests of the shading math the bounce loops call:
BSDF::{f, pdf, sample}, fresnelDielectric, AreaLight::{illuminance, emit, getRadiance}, the samplers and the camera
sample of SurfaceIntegrator::render — the device restatement (csrc/shading.cuh) in its host build against the
UNMODIFIED reference's functions (R/src/material/bsdf.cpp:24-335, fresnel.cpp:3-30, scene/light.cpp:4-100,
sampler/sampler.cpp:3-135), entry by entry on random + edge-case inputs (delta / glass / Phong materials, grazing
directions around the cmp() epsilon, lobe-selection edges, black materials).
Host build and reference share libm and IEEE no-FMA arithmetic, so the bar here is BIT equality; the CUDA build of the
same functions is checked in test_gpu_shading_kat.py (1e-5 relative: cosf / sinf / powf differ by ulps)."""
import numpy as np
import pytest

import shading_inputs as S
import util
from hostsim_py import HostSim


@pytest.mark.parametrize("what", sorted(S.NAMES))
def test_host_build_of_shading_equals_reference_bitwise(wrt, have_ref, what):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    sc = S.kat_scene()
    hs = util.host_scene(wrt, sc); sim = HostSim(hs.desc(), hs)
    ref = util.ref_scene(sc)
    inp = S.all_batches(sc, 20000)[what]
    a = ref.shading(what, inp, iparam=16)
    b = sim.debug_shading(what, inp, iparam=16, cam=hs.camera())
    same = (util.bits(a) == util.bits(b)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))
    bad = np.nonzero(~same.all(axis=1))[0]
    assert len(bad) == 0, "%s: %d of %d records differ, first: in=%s ref=%s ours=%s" % (S.NAMES[what], len(bad), len(inp), inp[bad[0]], a[bad[0]], b[bad[0]])
    assert np.abs(a).sum() > 0           # the batch exercises something


def test_kat_batches_cover_the_edge_cases(wrt, have_ref):
    """The BSDF batches must actually reach: invalid BSDFs (|cos wi| <= EPS), delta materials, every sampled lobe,
    black results, total internal reflection."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    sc = S.kat_scene()
    ref = util.ref_scene(sc)
    b = S.all_batches(sc, 20000)
    f = ref.shading(0, b[0])
    assert (f[:, 8] == 0).sum() > 50 and (f[:, 7] == 1).sum() > 1000 and (np.abs(f[:, :3]).sum(1) == 0).sum() > 1000
    s = ref.shading(1, b[1])
    types = set(int(t) for t in np.unique(s[:, 8]))
    assert {1, 2, 4, 8} <= types, types        # BSDF_REFLECTION, TRANSMISSION, DIFFUSE, GLOSSY were all sampled
    glass_inside = (b[1][:, 6] == 3) & (np.sum(b[1][:, 0:3] * b[1][:, 3:6], axis=1) < -0.05)
    assert ((s[:, 8] == 1) & glass_inside).sum() > 100        # reflection picked from inside glass (incl. total internal reflection)
