"""Diagnostic: C3 at 64 / 8 spp with different tail budgets — device time, rays handed over, extend launches and summed extend time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import wrt_b200 as W
import scenes, util
W.set_device(0)
sc = scenes.synthetic_torus_scene(n=708, width=1920, height=1080)
hs = util.host_scene(W, sc); scene = W.Scene(hs); cam = hs.camera()
film = torch.zeros((1080, 1920, 3), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for spp in (64, 8):
    p = W.PtParams(1920, 1080, spp, 5, 1000, 0, 1, 0.0)
    for budget in ("0", "200", "2000", "20000", "24"):
        os.environ["WRT_TAIL_BUDGET"] = budget
        for _ in range(2): scene.render_pt_dev(cam, p, film.data_ptr(), st)
        ms = []
        for _ in range(3):
            scene.reset_stats(); film.zero_(); scene.render_pt_dev(cam, p, film.data_ptr(), st); s = scene.stats(); ms.append(s.last_render_ms)
        print("spp %2d budget %6s: %8.2f ms  handed over %8d of %d rays  extend launches %3d  extend %.1f ms shade %.1f ms shadow %.1f ms  kernels %d" %
              (spp, budget, np.median(ms), s.suspended_rays, s.closest_rays, s.extend_launches, s.extend_ms, s.shade_ms, s.shadow_ms, s.kernel_launches), flush=True)
