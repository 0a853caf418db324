#!/usr/bin/env python
"""Headline benchmark: Mrays/s (and samples/s) of the path-tracing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|torus|cbox_dragon|c4]

A "step" is one full render of the workload: every closest-hit and shadow ray of the wavefront path
tracer (or BDPT for c4) over one frame of synthetic input.  Default workload = BASELINE.json
configs[2] ("c3"): synthetic 1 002 528-triangle displaced torus, 1920x1080 path tracing, 5 bounces,
64 spp — the largest single-GPU configuration the metric is quoted on (configs[0..1] are parity cases).

Printed JSON (one line, rank 0): value = whole-job Mrays/s with everything resident in HBM (CUDA
events, max over ranks); e2e = the same through the host-buffer C-ABI call (film copied back to pinned
host memory every step); roofline = the closest-hit (extend) kernel's algorithmic GB/s against the
measured HBM peak; cpu_baseline = the reference's own renderer timed on this box's host cores on a
bounded sample.  N > 1 (torchrun): every rank renders the full frame at 64 spp with its own RNG stream
(weak scaling), films are summed with one NCCL reduce inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    "c3": dict(desc="C3: synthetic 1,002,528-triangle displaced torus + 2-triangle area light, 1920x1080 PT, depth 5, 64 spp",
               integrator="pt", width=1920, height=1080, spp=64, depth=5, n=708),
    "torus": dict(desc="C1-class: torus.scene (13,486 triangles, glass + diffuse), 512x512 PT, depth 7, 256 spp",
                  integrator="pt", width=512, height=512, spp=256, depth=7, fixture="torus"),
    "cbox_dragon": dict(desc="C2-class: Cornell walls + dragon (2,584 triangles), 512x512 PT, depth 7, 256 spp",
                        integrator="pt", width=512, height=512, spp=256, depth=7, fixture="cbox_dragon"),
    "c5": dict(desc="C5: synthetic 10,008,338-triangle displaced torus + 100,000 spheres, 3840x2160 PT, depth 5, 16 spp per step (of 1024)",
               integrator="pt", width=3840, height=2160, spp=16, depth=5, n=2237, n_spheres=100000),
    "c5_small": dict(desc="C5 at 1/5 scale: 2,000,000-triangle displaced torus + 20,000 spheres, 3840x2160 PT, depth 5, 16 spp",
                     integrator="pt", width=3840, height=2160, spp=16, depth=5, n=1000, n_spheres=20000),
    "whitted_torus": dict(desc="Whitted (SURVEY 8(f)4): torus.scene (13,486 triangles, glass + diffuse), 512x512, depth 7, 256 spp",
                          integrator="whitted", width=512, height=512, spp=256, depth=7, fixture="torus"),
    "c4": dict(desc="C4: closed Cornell box + area light, BDPT 1440x1440, 16 iterations per step (of 256), controlLength 3",
               integrator="bdpt", width=1440, height=1440, iterations=16, n=0),
}


# DRAM bytes per ray of k_pt_extend from the committed ncu captures (profiles/r1_final_launches_and_ncu.md: C3, one
# 33 554 432-ray launch, dram__bytes_read.sum 4.749 GB + dram__bytes_write.sum 1.546 GB; profiles/r1_ncu_extend_c5.md: C5,
# one 33 554 432-ray launch, 22.52 GB + 4.88 GB).
NCU_DRAM_BYTES_PER_RAY = {"c3": (4.749402e9 + 1.546372e9) / 33554432.0, "c5": (22.523734e9 + 4.877666e9) / 33554432.0}
NCU_SOURCE = {"c3": "profiles/r1_final_launches_and_ncu.md", "c5": "profiles/r1_ncu_extend_c5.md"}


def make_scene(w):
    import scenes
    if "fixture" in w:
        sc, z = scenes.load_fixture(w["fixture"])
        return sc
    if w["integrator"] == "bdpt":
        return scenes.cornell_box_scene(w["width"], w["height"], closed=True)
    return scenes.synthetic_torus_scene(n=w["n"], width=w["width"], height=w["height"], n_spheres=w.get("n_spheres", 0))


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own renderer on host cores
# ---------------------------------------------------------------------------------------------------
def _ref_worker(conn, w, rank, nproc):
    """One reference process: builds the scene with the reference's own KD builder, then renders its
    share of a centre crop (1 spp) each time it is told to."""
    try:
        from oracle import refpy
        sc = make_scene(w)
        kind = w["integrator"]
        ref = refpy.RefScene(kind)
        t0 = time.time()
        ref.build(sc.materials, sc.kind, sc.data, sc.matid, sc.lights, sc.cam12, sc.width, sc.height)
        conn.send(("ready", time.time() - t0))
        stride = w["row_stride"]                              # every stride-th row of the whole frame
        my_rows = range(rank * stride, sc.height, nproc * stride)
        n_rows = len(my_rows)
        cw = sc.width
        while True:
            msg = conn.recv()
            if msg == "stop":
                break
            ref.reset_traverse_calls()
            t0 = time.perf_counter()
            if kind == "pt":
                if n_rows:
                    ref.render_pt_rows(1, w["depth"], 5489 + rank + 97 * msg, rank * stride, sc.height, 0, cw,
                                       want_film=False, row_stride=nproc * stride)
                samples = n_rows * cw
            elif kind == "whitted":
                ref.render_whitted(1, w["depth"], seed=5489 + rank + 97 * msg)     # whole frame per process
                samples = sc.width * sc.height
            else:
                ref.render_bdpt(1, seed=5489 + rank + 97 * msg)
                samples = sc.width * sc.height
            dt = time.perf_counter() - t0
            conn.send((ref.traverse_calls(), samples, dt))
    except Exception as e:  # pragma: no cover
        conn.send(("error", repr(e)))


def run_reference(w, steps, warmup, nproc):
    """Returns per-step (rays, samples, seconds) aggregated over nproc independent reference processes
    (the reference is single-threaded and not re-entrant: multi-core = independent processes)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    procs = []
    for r in range(nproc):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, w, r, nproc), daemon=True)
        p.start()
        procs.append((p, a))
    build_s = 0.0
    for p, a in procs:
        m = a.recv()
        if m[0] != "ready":
            raise RuntimeError("reference worker failed: %r" % (m,))
        build_s = max(build_s, m[1])
    out = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        for p, a in procs:
            a.send(s)
        res = [a.recv() for p, a in procs]
        wall = time.perf_counter() - t0
        if s >= warmup:
            out.append((sum(r[0] for r in res), sum(r[1] for r in res), wall))
    for p, a in procs:
        a.send("stop")
    return out, build_s


def reference_procs(w):
    """All the host threads the reference can use = independent processes, bounded by memory: the reference's
    KD build keeps every node's event lists (about 2 GB per million primitives)."""
    n = os.cpu_count() or 1
    try:
        import psutil
        per_proc_gb = 0.5 + 2.2 * (w.get("n", 0) ** 2 * 2 / 1e6 if w.get("n") else 0.05)
        n = min(n, max(1, int(psutil.virtual_memory().available / 2 ** 30 * 0.6 / per_proc_gb)))
    except Exception:
        n = min(n, 8)
    return max(1, min(n, 64))


def reference_workload(w):
    w = dict(w)
    if w["integrator"] == "pt":
        # bounded sample: 1 spp on every k-th row of the whole frame (same ray mix as the full frame)
        w["row_stride"] = 8 if w["width"] * w["height"] >= 2 ** 20 else 2
    elif w["integrator"] == "whitted":
        w["row_stride"] = 1              # whole frame at 1 spp per process
    else:
        w["width"] = w["height"] = 256   # BDPT must render whole (square) frames: 1 iteration at 256^2
        w["row_stride"] = 1
    return w


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("WRT_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (PT) / iterations (BDPT)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every GPU renders the workload's spp; strong: the spp (iterations) are sharded over the GPUs "
                         "(samples k = rank, rank + N, ... of the same stratification grid, SURVEY.md 8e)")
    args = ap.parse_args()

    w = dict(WORKLOADS[args.workload])
    if args.spp:
        w["iterations" if w["integrator"] == "bdpt" else "spp"] = args.spp
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric, unit = "path-tracing ray throughput (closest-hit + shadow rays)", "Mrays/s"
    if w["integrator"] == "bdpt":
        metric = "bidirectional path-tracing ray throughput (closest-hit + connection rays)"
    if w["integrator"] == "whitted":
        metric = "Whitted ray-tree throughput (closest-hit + occlusion rays)"

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import refpy
        if not refpy.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libwrt_ref.so was not built (no /root/reference at build time)"}))
            return 0
        nproc = reference_procs(w)
        rw = reference_workload(w)
        res, build_s = run_reference(rw, args.steps, max(args.warmup, 1), nproc)
        rays = sum(r[0] for r in res); samples = sum(r[1] for r in res); secs = sum(r[2] for r in res)
        v = rays / secs / 1e6
        if rw["integrator"] == "pt":
            sample = ("1 spp on every %d-th row of the %dx%d frame per step, rows dealt round-robin to %d independent reference processes"
                      % (rw["row_stride"], rw["width"], rw["height"], nproc))
        else:
            sample = ("1 %s of the whole %dx%d frame per step in each of %d independent reference processes"
                      % ("iteration" if rw["integrator"] == "bdpt" else "spp", rw["width"], rw["height"], nproc))
        line = {"metric": metric, "value": v, "unit": unit, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(len(res), 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["desc"], "sample": sample},
                "samples_per_s": samples / secs, "rays_per_sample": rays / max(samples, 1),
                "cpu_baseline": {"value": v, "unit": unit, "cores": nproc, "kind": "reference", "sample": sample,
                                 "kd_build_s": build_s},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- B200 arm
    import torch
    import wrt_b200 as W
    import util
    if not torch.cuda.is_available() or W.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    W.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t0 = time.time()
    sc = make_scene(w)
    hs = util.host_scene(W, sc)
    kd_build_s = time.time() - t0
    t0 = time.time()
    scene = W.Scene(hs)
    upload_s = time.time() - t0
    cam = hs.camera()
    npix = w["width"] * w["height"]

    if w["integrator"] in ("pt", "whitted"):
        spp = w["spp"]
        whitted = w["integrator"] == "whitted"

        strong = args.scaling == "strong" and world > 1
        if strong and spp < world:
            raise SystemExit("bench.py: --scaling strong needs at least one sample per pixel per GPU")

        def params(scale):
            if strong:      # one image: rank g renders samples g, g + N, ... of the shared grid with the shared seed
                return W.shard_pt(W.PtParams(w["width"], w["height"], spp, w["depth"], 1000, 0, 1, 0.0), rank, world)
            return W.PtParams(w["width"], w["height"], spp, w["depth"], 1000 + rank, 0, 1, scale)

        def render_dev(film):
            film.zero_()
            fn = scene.render_whitted_dev if whitted else scene.render_pt_dev
            fn(cam, params(1.0 / (spp * world)), film.data_ptr(), torch.cuda.current_stream().cuda_stream)

        def render_host(buf):
            (scene.render_whitted if whitted else scene.render_pt)(cam, params(0.0), buf)
        samples_per_step = npix * spp if not strong else npix * spp / world
    else:
        iters = w["iterations"]

        strong = args.scaling == "strong" and world > 1
        if strong and iters < world:
            raise SystemExit("bench.py: --scaling strong needs at least one iteration per GPU")

        def params(scale):
            if strong:
                return W.shard_bdpt(W.BdptParams(w["width"], w["height"], iters, 0, 10, 3, 1000, 0, 1, 0.0, 0), rank, world)
            return W.BdptParams(w["width"], w["height"], iters, 0, 10, 3, 1000 + rank, 0, 1, scale, 0)

        def render_dev(film):
            film.zero_()
            scene.render_bdpt_dev(cam, params(1.0 / (iters * world)), film.data_ptr(), torch.cuda.current_stream().cuda_stream)

        def render_host(buf):
            scene.render_bdpt(cam, params(0.0), buf)
        samples_per_step = npix * iters if not strong else npix * iters / world

    film = torch.zeros((w["height"], w["width"], 3), dtype=torch.float32, device="cuda")
    host_film_t = torch.empty((w["height"], w["width"], 3), dtype=torch.float32, pin_memory=True)
    host_film = host_film_t.numpy()

    # reference-semantics work per ray of THIS ray mix (1 spp / 1 iteration counting render): B_ray
    b_ray, visits, b_ray_k, visits_k = None, None, None, None
    if rank == 0 and w["integrator"] == "pt":
        scene.set_counting(True); scene.reset_stats()
        scene.render_pt(cam, W.PtParams(w["width"], w["height"], 1, w["depth"], 7, 0, 1, 0.0), host_film)
        s = scene.stats(); scene.set_counting(False)
        nr = float(s.closest_rays + s.shadow_rays)
        visits = {"inner": s.inner_visits / nr, "leaf": s.leaf_visits / nr, "tri": s.tri_tests / nr, "sphere": s.sphere_tests / nr}
        b_ray = 40 + 8 * visits["inner"] + 8 * visits["leaf"] + 40 * visits["tri"] + 20 * visits["sphere"]
        # the same accounting for the PRUNED traversal the timed kernels run (their own work, 32-byte nodes, 48-byte records)
        scene.set_counting(2); scene.reset_stats()
        scene.render_pt(cam, W.PtParams(w["width"], w["height"], 1, w["depth"], 7, 0, 1, 0.0), host_film)
        s = scene.stats(); scene.set_counting(False)
        nr = float(s.closest_rays + s.shadow_rays)
        visits_k = {"inner": s.inner_visits / nr, "leaf": s.leaf_visits / nr, "tri": s.tri_tests / nr, "sphere": s.sphere_tests / nr}
        b_ray_k = 40 + 32 * (visits_k["inner"] + visits_k["leaf"]) + 48 * (visits_k["tri"] + visits_k["sphere"])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        render_dev(film)
        if dist is not None:
            dist.reduce(film, 0)
    barrier()
    scene.reset_stats()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ext_ms = ext_rays = ext_launches = 0.0
    barrier()
    e0.record()
    for _ in range(args.steps):
        render_dev(film)
        if dist is not None:
            dist.reduce(film, 0)
        st = scene.stats()
        ext_ms += st.extend_ms; ext_rays += st.extend_rays; ext_launches += st.extend_launches
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    st = scene.stats()
    rays = float(st.closest_rays + st.shadow_rays)
    launches = int(st.kernel_launches)
    mean_radiance = float(torch.nanmean(film).item()) if rank == 0 else 0.0   # Whitted films hold the reference's NaN pixels
    if dist is not None:
        tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        rr = torch.tensor([rays, float(launches)], dtype=torch.float64, device="cuda"); dist.all_reduce(rr)
        rays, launches = float(rr[0].item()), int(rr[1].item())

    # end to end through the host-buffer C-ABI call (film copied to pinned host memory every step)
    for _ in range(1):
        render_host(host_film)
    barrier()
    t0 = time.perf_counter()
    scene.reset_stats()
    for _ in range(args.steps):
        render_host(host_film)
        if dist is not None:   # host path at N>1: stage through the device film of rank 0
            film.copy_(host_film_t, non_blocking=True); film.mul_(1.0 if strong else 1.0 / world); dist.reduce(film, 0)
            if rank == 0:
                host_film_t.copy_(film)
    barrier()
    e2e_s = time.perf_counter() - t0
    st2 = scene.stats()
    e2e_rays = float(st2.closest_rays + st2.shadow_rays)
    if dist is not None:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_s = float(tt.item())
        rr = torch.tensor([e2e_rays], dtype=torch.float64, device="cuda"); dist.all_reduce(rr); e2e_rays = float(rr.item())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = rays / ms / 1e3
    total_samples = samples_per_step * args.steps * world
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "integrator": w["integrator"], "traversal": "pruned (bit-exact vs exact, tests/test_gpu_traversal.py)",
                   "per_gpu": ("full frame, %d %s per GPU, disjoint RNG streams; films summed by one NCCL reduce" if not strong else
                               "full frame, %d %s in total, dealt round-robin to the GPUs (same grid and seed as 1 GPU); films summed by one NCCL reduce") %
                              (w.get("spp", w.get("iterations")), "iterations" if w["integrator"] == "bdpt" else "spp"),
                   "l2": "inputs larger than L2: the path pool (2^24 slots x 176 B in 2 concurrent sub-pools) is rewritten every bounce; the scene is meant to stay L2-resident",
                   "prims": int(sc.n_prims), "kd_build_s": round(kd_build_s, 2), "upload_s": round(upload_s, 2)},
        "samples_per_s": total_samples / (ms / 1e3), "rays_per_sample": rays / total_samples,
        "mean_radiance": mean_radiance,
        "gpu_launches": launches,
        "clocks": clk,
        "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": unit,
                "h2d_bytes_per_step": int(C_sizeof_inputs(W, w)), "d2h_bytes_per_step": int(npix * 12),
                "ms_per_step": 1e3 * e2e_s / args.steps},
    }
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    if b_ray is not None and ext_ms > 0:
        # dominant kernel = closest-hit traversal (k_pt_extend): algorithmic bytes of the reference-semantics
        # traversal of the rays it traced / its own CUDA-event time
        achieved = (ext_rays * b_ray) / (ext_ms * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "k_pt_extend<pruned>", "achieved": achieved, "peak": peak,
                            "note": "launch durations are CUDA-event times of launches that overlap with the other sub-pool's kernels (2 streams), so this is a lower bound of the kernel's stand-alone rate",
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                            "unit": "GB/s", "frac": achieved / peak,
                            "traffic": (NCU_DRAM_BYTES_PER_RAY.get(args.workload) or 0) * ext_rays / max(ext_launches, 1) or None,
                            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per ray of one ncu --set full capture "
                                              "(%s) x rays per launch" % NCU_SOURCE.get(args.workload, "none for this workload"),
                            "frac_note": "frac uses the REFERENCE-semantics bytes SURVEY 8(d) defines (full traversal, no early exit); PRUNED traversal "
                                         "skips most of that work, so frac > 1 is possible and says how much of the reference's traffic is avoided, not a "
                                         "bandwidth; frac_own_work = the kernels' own algorithmic bytes / peak.  The scene is L2-resident and the "
                                         "kernel is bound by issue slots and L2 latency (profiles/): DRAM traffic per launch is in `traffic`",
                            "frac_own_work": (ext_rays * b_ray_k) / (ext_ms * 1e-3) / 1e9 / peak,
                            "bytes_per_ray": b_ray, "visits_per_ray_reference_semantics": visits,
                            "kernel_own_work": {"visits_per_ray": visits_k, "bytes_per_ray": b_ray_k,
                                                "achieved_GBps": (ext_rays * b_ray_k) / (ext_ms * 1e-3) / 1e9,
                                                "note": "what the PRUNED kernels actually fetch (not counting skipped nodes): "
                                                        "pruning removes most of the reference-semantics bytes, which is why frac can exceed 1"},
                            "avg_launch_ms": ext_ms / max(ext_launches, 1), "launches": int(ext_launches),
                            "rays_per_launch": ext_rays / max(ext_launches, 1),
                            "kernel_share_of_step": ext_ms / ms,
                            "extend_mrays_per_s": ext_rays / ext_ms / 1e3}
    else:
        line["roofline"] = None
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w, unit)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def C_sizeof_inputs(W, w):
    import ctypes
    return ctypes.sizeof(W.Camera) + ctypes.sizeof(W.BdptParams if w["integrator"] == "bdpt" else W.PtParams)


def cpu_baseline(w, unit):
    """The reference renderer (oracle/_ref) on ONE host core, bounded sample of the same workload."""
    from oracle import refpy
    if not refpy.available():
        return {"value": None, "unit": unit, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    rw = reference_workload(w)
    if rw["integrator"] == "pt":
        rw["row_stride"] *= 2
    res, build_s = run_reference(rw, 1, 0, 1)
    rays, samples, secs = res[0]
    return {"value": rays / secs / 1e6, "unit": unit, "cores": 1, "kind": "reference",
            "sample": "1 spp on every %d-th row of the %dx%d frame (%d rays, %.1f s); reference KD build %.1f s not timed"
                      % (rw["row_stride"], rw["width"], rw["height"], rays, secs, build_s),
            "samples_per_s": samples / secs, "rays_per_sample": rays / max(samples, 1)}


if __name__ == "__main__":
    sys.exit(main())
