"""Quick device-side throughput probe (not the bench): Mrays/s of the closest-hit kernel on primary and
secondary batches, both traversal modes, plus a PT render rate."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import wrt_b200 as W
import scenes, util

def probe(name, sc, cam, w, h, spp=4, depth=5):
    t0 = time.time(); hs = util.host_scene(W, sc); tb = time.time() - t0
    scene = W.Scene(hs)
    rays = W.generate_rays(cam or hs.camera(), scenes.pixel_centres(w, h))
    a = scene.intersect(rays, full=True)
    hit = a[0] >= 0
    r2 = W.make_rays(scenes.bounce_rays(a[2], a[3], hit))
    for label, r in (("primary", rays), ("secondary", r2)):
        d_rays = torch.from_numpy(r).cuda(); n = len(r)
        d_prim = torch.empty(n, dtype=torch.int32, device="cuda"); d_t = torch.empty(n, dtype=torch.float32, device="cuda")
        for mode, mname in ((W.TRAVERSE_EXACT, "exact"), (W.TRAVERSE_PRUNED, "pruned")):
            scene.set_traversal(mode)
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(2): scene.intersect_dev(d_rays.data_ptr(), n, d_prim.data_ptr(), d_t.data_ptr(), st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): scene.intersect_dev(d_rays.data_ptr(), n, d_prim.data_ptr(), d_t.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("%-14s %-9s %-6s n=%8d  %8.3f ms  %9.1f Mrays/s" % (name, label, mname, n, ms, n / ms / 1e3), flush=True)
    cv = scene.count_visits(rays)
    print("%-14s visits/ray (reference semantics, primary): inner %.1f leaf %.1f tri %.1f sph %.1f" %
          (name, cv["inner"] / cv["rays"], cv["leaf"] / cv["rays"], cv["tri"] / cv["rays"], cv["sphere"] / cv["rays"]))
    scene.set_traversal(W.TRAVERSE_PRUNED)
    p = W.PtParams(w, h, spp, depth, 1, 0, 1, 0.0)
    scene.render_pt(hs.camera() if cam is None else cam, p)
    scene.reset_stats()
    film = scene.render_pt(hs.camera() if cam is None else cam, p)
    s = scene.stats()
    nr = s.closest_rays + s.shadow_rays
    print("%-14s PT %dx%d spp %d depth %d: %.1f ms, %.2f Mrays/s, %.2f Msamples/s, rays/sample %.2f, launches %d, mean %.4f (kd build %.1fs)"
          % (name, w, h, spp, depth, s.last_render_ms, nr / s.last_render_ms / 1e3, s.samples / s.last_render_ms / 1e3,
             nr / s.samples, s.kernel_launches, film.mean(), tb), flush=True)

if __name__ == "__main__":
    W.set_device(0)
    sc, z = scenes.load_fixture("torus")
    probe("torus", sc, W.Camera.from_ref_array(z["cam45"]), 512, 512, spp=16, depth=7)
    sc, z = scenes.load_fixture("cbox_dragon")
    probe("cbox_dragon", sc, W.Camera.from_ref_array(z["cam45"]), 512, 512, spp=16, depth=7)
    if "--big" in sys.argv:
        sc = scenes.synthetic_torus_scene(n=708, width=1920, height=1080)
        probe("synthetic_1m", sc, None, 1920, 1080, spp=4, depth=5)
