"""Which rays make the tail of an extend launch?  C3 scene: camera rays (jittered) and bounce rays in chunks of 64 K through
wrt_trace_closest_dev, chunk times -> slowest chunks -> bisection by time down to single rays; prints the rays, their visit counts
(PRUNED and EXACT counting modes) and how long one such ray takes alone."""
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
scene.set_traversal(W.TRAVERSE_PRUNED)
rng = np.random.default_rng(11)
px = scenes.pixel_centres(1920, 1080)
cams = [W.generate_rays(cam, px + rng.uniform(-0.5, 0.5, px.shape).astype(np.float32)) for _ in range(6)]
rays = np.concatenate(cams)
a = scene.intersect(cams[0], full=True); hit = a[0] >= 0
b1 = W.make_rays(scenes.bounce_rays(a[2], a[3], hit))
a2 = scene.intersect(b1, full=True); hit2 = a2[0] >= 0
b2 = W.make_rays(scenes.bounce_rays(a2[2], a2[3], hit2))
st = torch.cuda.current_stream().cuda_stream

def time_rays(r, reps=2):
    d = torch.from_numpy(np.ascontiguousarray(r)).cuda(); n = len(r)
    dp = torch.empty(n, dtype=torch.int32, device="cuda"); dt = torch.empty(n, dtype=torch.float32, device="cuda")
    scene.intersect_dev(d.data_ptr(), n, dp.data_ptr(), dt.data_ptr(), st); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); scene.intersect_dev(d.data_ptr(), n, dp.data_ptr(), dt.data_ptr(), st); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def visits(r, mode):
    scene.set_counting(mode); scene.reset_stats(); v = scene.count_visits(r); scene.set_counting(False); return v

for label, R in (("camera", rays), ("bounce1", b1), ("bounce2", b2)):
    B = 65536
    t = np.array([time_rays(R[i:i + B]) for i in range(0, len(R), B)])
    print("%s: %d rays in %d chunks: chunk time median %.3f ms, p90 %.3f, max %.3f" % (label, len(R), len(t), np.median(t), np.percentile(t, 90), t.max()), flush=True)
    for w in np.argsort(-t)[:3]:
        lo, hi = w * B, min((w + 1) * B, len(R))
        while hi - lo > 1:
            mid = (lo + hi) // 2
            ta, tb = time_rays(R[lo:mid]), time_rays(R[mid:hi])
            if ta >= tb: hi = mid
            else: lo = mid
        r = R[lo:lo + 1]
        vp, ve = visits(r, 2), visits(r, True)
        res = scene.intersect(r)
        print("  chunk %.3f ms -> ray alone %.3f ms: o=%s d=%s | pruned inner %d leaf %d tri %d | exact inner %d leaf %d tri %d | prim %d t %.4f"
              % (t[w], time_rays(r), np.array2string(r[0, :3], precision=5), np.array2string(r[0, 3:6], precision=7), vp["inner"], vp["leaf"], vp["tri"],
                 ve["inner"], ve["leaf"], ve["tri"], res[0][0], res[1][0]), flush=True)
