"""How much does the ORDER of the rays in a batch matter to the pooled closest-hit kernel?  (binning of
secondary rays.)  Same rays, same kernel (wrt_trace_closest on device pointers, PRUNED), four orders: as generated (camera rays:
scanline; bounce rays: order of the pixels they came from), randomly permuted, sorted by (origin cell 32^3 Morton, direction octant).
Results are identical by construction (asserted)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import wrt_b200 as W
import scenes, util


def morton3(c):
    def spread(v):
        v = v.astype(np.uint64) & 0x3ff
        v = (v | (v << 16)) & 0x30000ff
        v = (v | (v << 8)) & 0x300f00f
        v = (v | (v << 4)) & 0x30c30c3
        v = (v | (v << 2)) & 0x9249249
        return v
    return spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)


def sort_key(rays, lo, hi, cells=32):
    o = rays[:, 0:3].astype(np.float64); d = rays[:, 3:6]
    c = np.clip(((o - lo) / (hi - lo) * cells).astype(np.int64), 0, cells - 1)
    octant = (d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2)
    return (morton3(c) << 3) | octant


def time_batch(scene, r, reps=5):
    d_rays = torch.from_numpy(np.ascontiguousarray(r)).cuda(); n = len(r)
    d_prim = torch.empty(n, dtype=torch.int32, device="cuda"); d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): scene.intersect_dev(d_rays.data_ptr(), n, d_prim.data_ptr(), d_t.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): scene.intersect_dev(d_rays.data_ptr(), n, d_prim.data_ptr(), d_t.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, d_prim.cpu().numpy(), d_t.cpu().numpy()


if __name__ == "__main__":
    W.set_device(0)
    n = int(os.environ.get("PROBE_N", "708"))
    sc = scenes.synthetic_torus_scene(n=n, width=1920, height=1080)
    hs = util.host_scene(W, sc); scene = W.Scene(hs)
    scene.set_traversal(W.TRAVERSE_PRUNED)
    rays = W.generate_rays(hs.camera(), scenes.pixel_centres(1920, 1080))
    rays = np.concatenate([rays] * 4)           # 8.3 M rays: four passes over the frame, like 4 spp of regenerated camera rays
    a = scene.intersect(rays[: len(rays) // 4], full=True)
    hit = a[0] >= 0
    r2 = W.make_rays(scenes.bounce_rays(a[2], a[3], hit))
    r2 = np.concatenate([r2] * 4)
    pts = sc.data.reshape(-1, 3) if hasattr(sc, "data") else rays[:, :3]
    lo = np.minimum(r2[:, :3].min(0), rays[:, :3].min(0)) - 1e-3; hi = np.maximum(r2[:, :3].max(0), rays[:, :3].max(0)) + 1e-3
    rng = np.random.default_rng(1)
    for label, r in (("camera", rays), ("bounce", r2)):
        base = None
        orders = [("as generated", np.arange(len(r))), ("shuffled", rng.permutation(len(r))), ("sorted cell x octant", np.argsort(sort_key(r, lo, hi), kind="stable"))]
        for oname, perm in orders:
            ms, prim, t = time_batch(scene, r[perm])
            inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
            if base is None: base = (prim[inv], t[inv])
            else: assert np.array_equal(prim[inv], base[0]) and np.array_equal(t[inv].view(np.uint32), base[1].view(np.uint32))
            print("%-7s %-22s n=%9d  %8.3f ms  %8.1f Mrays/s" % (label, oname, len(r), ms, len(r) / ms / 1e3), flush=True)
