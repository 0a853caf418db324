"""TEST INFRASTRUCTURE — small committed RNG tapes recorded from the UNMODIFIED reference (oracle/_ref), with the films the
reference rendered from them.  tests/test_tape_parity.py replays them through the device code (hostsim on the CPU, CUDA on
the GPU box, where /root/reference does not exist) and compares per pixel.

    python tests/golden/make_golden_tapes.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import scenes  # noqa: E402
import util  # noqa: E402

RES, STRIDE = 24, 64
out_dir = os.path.dirname(os.path.abspath(__file__))
sc = scenes.small_mixed_scene(RES, RES)
ref = util.ref_scene(sc, "pt")
film, tape, rgb, draws = ref.render_pt_tape(1, 5, seed=5489, stride=STRIDE)
np.savez_compressed(os.path.join(out_dir, "tape_pt_small_mixed.npz"), film=film, tape=tape, draws=draws, res=RES, spp=1, depth=5, stride=STRIDE)
print("pt: mean %.5f, max draws %d" % (film.mean(), draws.max()))
ref = util.ref_scene(sc, "bdpt")
film, tape, draws = ref.render_bdpt_tape(1, seed=5489, stride=STRIDE)
np.savez_compressed(os.path.join(out_dir, "tape_bdpt_small_mixed.npz"), film=film, tape=tape, draws=draws, res=RES, iterations=1, stride=STRIDE)
print("bdpt: mean %.5f, max draws %d" % (film.mean(), draws.max()))
