"""Time of single long rays (a tail of an extend launch in isolation): C3 scene, the camera / bounce rays long_ray_probe.py found."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import wrt_b200 as W
import scenes, util
W.set_device(0)
sc = scenes.synthetic_torus_scene(n=708, width=1920, height=1080)
hs = util.host_scene(W, sc); scene = W.Scene(hs)
scene.set_traversal(W.TRAVERSE_PRUNED)
R = np.array([[2.6, -2.6, 1.7, -0.6839116, 0.6743881, -0.2783267, 0, 1e7],
              [2.6, -2.6, 1.7, -0.6835189, 0.6748181, -0.2782488, 0, 1e7],
              [-0.6501, 0.60616, 0.37675, -0.4000501, 0.2225898, 0.8890522, 0, 1e7],
              [-0.06194, -0.3232, 2.19914, 0.3575231, 0.4105029, -0.8388472, 0, 1e7]], np.float32)
st = torch.cuda.current_stream().cuda_stream
for i in range(len(R)):
    r = W.make_rays(R[i:i + 1, :6]) if hasattr(W, "make_rays") else R[i:i + 1]
    d = torch.from_numpy(np.ascontiguousarray(r)).cuda()
    dp = torch.empty(1, dtype=torch.int32, device="cuda"); dt = torch.empty(1, dtype=torch.float32, device="cuda")
    for _ in range(2): scene.intersect_dev(d.data_ptr(), 1, dp.data_ptr(), dt.data_ptr(), st)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); scene.intersect_dev(d.data_ptr(), 1, dp.data_ptr(), dt.data_ptr(), st); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("ray %d alone: %.3f ms (prim %d, t %.4f)" % (i, min(ts), int(dp.item()), float(dt.item())), flush=True)
