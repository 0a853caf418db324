"""C1 (torus.scene as shipped): what the longest camera / bounce rays look like (chunk times -> bisection), and the leaf-size histogram of the tree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import wrt_b200 as W
import scenes, util
W.set_device(0)
sc, z = scenes.load_fixture("torus")
hs = util.host_scene(W, sc); scene = W.Scene(hs); cam = W.Camera.from_ref_array(z["cam45"])
a_ = hs.arrays(); tr = a_["tree"]
nref = np.asarray(tr["nref"]); axis = np.asarray(tr["axis"])
leaf = nref[axis < 0]
print("tree: %d nodes, %d leaves, refs %d; leaf size median %d p90 %d p99 %d max %d; leaves >= 64 refs: %d holding %d refs" %
      (len(axis), len(leaf), leaf.sum(), np.median(leaf), np.percentile(leaf, 90), np.percentile(leaf, 99), leaf.max(), (leaf >= 64).sum(), leaf[leaf >= 64].sum()))
scene.set_traversal(W.TRAVERSE_PRUNED)
rng = np.random.default_rng(5)
px = scenes.pixel_centres(512, 512)
rays = W.generate_rays(cam, px + rng.uniform(-0.5, 0.5, px.shape).astype(np.float32))
a = scene.intersect(rays, full=True); hit = a[0] >= 0
b1 = W.make_rays(scenes.bounce_rays(a[2], a[3], hit))
st = torch.cuda.current_stream().cuda_stream
def time_rays(r, reps=3):
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
for label, R in (("camera", rays), ("bounce1", b1)):
    print("%s: all %d rays in one launch: %.3f ms" % (label, len(R), time_rays(R)))
    B = 8192
    t = np.array([time_rays(R[i:i + B]) for i in range(0, len(R), B)])
    print("  chunks of %d: median %.3f ms, p90 %.3f, max %.3f" % (B, np.median(t), np.percentile(t, 90), t.max()), flush=True)
    for w in np.argsort(-t)[:3]:
        lo, hi = w * B, min((w + 1) * B, len(R))
        while hi - lo > 1:
            mid = (lo + hi) // 2
            ta, tb = time_rays(R[lo:mid]), time_rays(R[mid:hi])
            if ta >= tb: hi = mid
            else: lo = mid
        r = R[lo:lo + 1]
        ve = visits(r, True)
        res = scene.intersect(r)
        print("  chunk %.3f ms -> ray alone %.3f ms: o=%s d=%s | exact inner %d leaf %d tri %d | prim %d t %.4f"
              % (t[w], time_rays(r), np.array2string(r[0, :3], precision=4), np.array2string(r[0, 3:6], precision=6), ve["inner"], ve["leaf"], ve["tri"], res[0][0], res[1][0]), flush=True)
