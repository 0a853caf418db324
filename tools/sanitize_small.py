"""Tiny end-to-end pass for compute-sanitizer (T5): trace queries, PT and BDPT renders on a mixed scene."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import wrt_b200 as W
import scenes, util, engines
W.set_device(0)
sc = scenes.small_mixed_scene(32, 32)
hs = util.host_scene(W, sc); scene = W.Scene(hs); cam = hs.camera()
rays = np.concatenate([W.generate_rays(cam, scenes.pixel_centres(32, 32)), W.make_rays(engines.adversarial_rays(sc, 2000))])
for mode in (W.TRAVERSE_EXACT, W.TRAVERSE_PRUNED):
    scene.set_traversal(mode)
    a = scene.intersect(rays, full=True)
    q = scenes.nee_queries(a[2], (a[0] >= 0) & (a[5] > 0), sc.lights)
    scene.occluded(q); scene.intersect_any(rays); scene.shadowRayTest(rays[:100], a[2][:100]); scene.count_visits(rays)
f1 = scene.render_pt(cam, W.PtParams(32, 32, 4, 5, 1, 0, 1, 0.0))
f2 = scene.render_bdpt(cam, W.BdptParams(32, 32, 2, 0, 10, 3, 1, 0, 1, 0.0, 1))
print("sanitize_small ok", float(f1.mean()), float(f2.mean()))
