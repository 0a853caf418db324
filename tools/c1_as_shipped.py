"""C1 as shipped (BASELINE configs[0]): torus.scene, 512 x 512, 1 spp, MAX_TRACING_DEPTH 7 — device time of one wrt_render_pt call
and the reference's own CPU time for the same frame (oracle/_ref, one core), on whatever box this runs on."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import wrt_b200 as W
import scenes, util
W.set_device(0)
sc, z = scenes.load_fixture("torus")
hs = util.host_scene(W, sc); scene = W.Scene(hs); cam = W.Camera.from_ref_array(z["cam45"])
p = W.PtParams(512, 512, 1, 7, 0, 0, 1, 0.0)
for _ in range(3): film = scene.render_pt(cam, p)
ms = []
for _ in range(10):
    t0 = time.perf_counter(); film = scene.render_pt(cam, p); wall = (time.perf_counter() - t0) * 1e3
    ms.append((scene.stats().last_render_ms, wall))
st = scene.stats()
print("stages of the last frame (summed CUDA-event durations): extend %.3f ms, shade %.3f ms, shadow %.3f ms over %d iterations with work; %d closest + %d shadow rays"
      % (st.extend_ms, st.shade_ms, st.shadow_ms, st.extend_launches, st.extend_rays, 0))
print("GPU: device %.3f ms, host wall incl. film copy %.3f ms (median of 10), mean radiance %.5f" % (np.median([m[0] for m in ms]), np.median([m[1] for m in ms]), film.mean()))
from oracle import refpy
if refpy.available():
    ref = util.ref_scene(sc)
    t0 = time.perf_counter(); rf = ref.render_pt(1, 7, seed=5489); print("reference, one core: %.2f s, mean radiance %.5f" % (time.perf_counter() - t0, rf.mean()))
