"""TEST INFRASTRUCTURE — converged reference films for the per-pixel image-parity tests.

    python tests/golden/make_golden_films.py [case ...]

Renders each case with the UNMODIFIED reference (oracle/_ref/libwrt_ref.so) at a high sample count, split over
processes x seeds (the reference is single-threaded), and stores the two half-averages `a` and `b` (disjoint seeds)
so a test can print the reference's own noise floor next to the device-vs-reference figure.  Needs /root/reference
(or a built oracle/_ref) — run here in the build container; the .npz files are committed.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# case -> (scene factory name, resolution, integrator, depth, samples per pixel (PT) / iterations (BDPT) per seed, seeds)
CASES = {
    "cornell_pt": ("cornell", 64, "pt", 5, 1024, 16),          # 16 384 spp
    "small_mixed_pt": ("small_mixed", 64, "pt", 5, 4096, 16),  # 65 536 spp
    "cornell_bdpt": ("cornell", 48, "bdpt", 0, 2048, 32),      # 65 536 iterations
    "small_mixed_bdpt": ("small_mixed", 48, "bdpt", 0, 2048, 48),   # seeds 3000..: with seeds 1000.. one light-tracing vertex lands in the
                                                                     # camera plane and the reference aborts in Transform::tPoint's assert(wp != 0)
    "torus_pt": ("torus", 32, "pt", 7, 16384, 16),             # 262 144 spp: glass caustics, see the test's note
}


def make_scene(name, res):
    import scenes
    if name == "cornell":
        return scenes.cornell_box_scene(res, res)
    if name == "small_mixed":
        return scenes.small_mixed_scene(res, res)
    sc, z = scenes.load_fixture(name)
    sc.cam12 = sc.cam12.copy(); sc.cam12[9] = res; sc.cam12[10] = res; sc.width = sc.height = res
    return sc


def work(args):
    case, seed = args
    import util
    name, res, integ, depth, n, seeds = CASES[case]
    sc = make_scene(name, res)
    ref = util.ref_scene(sc, integ)
    base = 3000 if case == "small_mixed_bdpt" else 1000
    if integ == "pt":
        return ref.render_pt(n, depth, seed=base + seed).astype(np.float64)
    return ref.render_bdpt(n, seed=base + seed).astype(np.float64) / n


def main():
    cases = sys.argv[1:] or list(CASES)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        for case in cases:
            t0 = time.time()
            name, res, integ, depth, n, seeds = CASES[case]
            films = pool.map(work, [(case, s) for s in range(seeds)])
            a = np.mean(films[: seeds // 2], axis=0); b = np.mean(films[seeds // 2:], axis=0)
            m = 0.5 * (a + b)
            floor = np.sqrt(np.mean((a - b) ** 2)) / m.mean()
            np.savez_compressed(os.path.join(out_dir, "film_%s.npz" % case), a=a.astype(np.float32), b=b.astype(np.float32),
                                per_half=np.int64(n * seeds // 2), scene=name, res=res, integrator=integ, depth=depth)
            print("%s: %d x %d, %d samples per half, mean %.5f, two-half per-pixel rRMSE %.4f (the mean of both halves is "
                  "a factor 2 closer to the truth), %.0f s" % (case, res, res, n * seeds // 2, m.mean(), floor, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
