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
bounded sample.  N > 1 (torchrun): STRONG scaling by default — the workload's samples per pixel (iterations for BDPT)
are dealt round-robin to the ranks (same stratification grid, same RNG keys as the 1-GPU render), films
are summed with one NCCL reduce inside the timed region (its device time is reported as reduce_ms), and rank 0 checks
the N-GPU film against the 1-GPU film (T4: max relative difference).  --scaling weak: every rank renders the full spp.
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
    "c1": dict(desc="C1 (BASELINE configs[0]): torus.scene as shipped (13,486 triangles, glass + diffuse), 512x512 PT, depth 7, 1 spp",
               integrator="pt", width=512, height=512, spp=1, depth=7, fixture="torus"),
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


def ncu_metrics(workload):
    """Counters of the dominant kernel from the committed `ncu --set full` capture of this workload (profiles/ncu_metrics.json,
    written by tools/ncu_summary.py from the .ncu-rep of the same command): DRAM bytes per ray, issue-slot utilisation,
    threads per instruction.  They explain the live number; they are never measured under the profiler in this run."""
    try:
        m = json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics.json")))
        return m.get(workload)
    except Exception:
        return None


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
        w["width"] = w["height"] = 512   # BidirPathTracing::runIteration renders whole square frames (all W*H light paths, then all
        w["row_stride"] = 1              # camera paths): 1 iteration of the same scene at 512^2 (same path-length mix) instead of 1440^2
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
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="strong (default at N > 1): the spp (iterations) are sharded over the GPUs (samples k = rank, rank + N, ... of the "
                         "same stratification grid, SURVEY.md 8e); weak: every GPU renders the workload's spp")
    args = ap.parse_args()

    w = dict(WORKLOADS[args.workload])
    if args.spp:
        w["iterations" if w["integrator"] == "bdpt" else "spp"] = args.spp
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.scaling is None:
        n_units = w["iterations"] if w["integrator"] == "bdpt" else w["spp"]
        args.scaling = "strong" if n_units >= world else "weak"     # (C1 has 1 spp: it cannot be sharded)
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
    scene_gen_s = time.time() - t0          # synthetic geometry (numpy) or fixture load: not part of the library
    t0 = time.time()
    hs = util.host_scene(W, sc)
    kd_build_s = time.time() - t0           # HostScene from arrays incl. the exact multi-threaded KD build (host/kd_build.cpp)
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
    if rank == 0:
        def count_render(mode):
            scene.set_counting(mode); scene.reset_stats()
            if w["integrator"] == "bdpt":
                scene.render_bdpt(cam, W.BdptParams(w["width"], w["height"], 1, 0, 10, 3, 7, 0, 1, 0.0, 0), host_film)
            elif w["integrator"] == "whitted":
                scene.render_whitted(cam, W.PtParams(w["width"], w["height"], 1, w["depth"], 7, 0, 1, 0.0), host_film)
            else:
                scene.render_pt(cam, W.PtParams(w["width"], w["height"], 1, w["depth"], 7, 0, 1, 0.0), host_film)
            s_ = scene.stats(); scene.set_counting(False)
            nr = float(s_.closest_rays + s_.shadow_rays)
            return {"inner": s_.inner_visits / nr, "leaf": s_.leaf_visits / nr, "tri": s_.tri_tests / nr, "sphere": s_.sphere_tests / nr}
        visits = count_render(True)
        b_ray = 40 + 8 * visits["inner"] + 8 * visits["leaf"] + 40 * visits["tri"] + 20 * visits["sphere"]
        # the same accounting for the PRUNED traversal the timed kernels run (their own work, 32-byte nodes, 48-byte records)
        visits_k = count_render(2)
        b_ray_k = 40 + 32 * (visits_k["inner"] + visits_k["leaf"]) + 48 * (visits_k["tri"] + visits_k["sphere"])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    red0, red1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reduce_ms = 0.0

    def reduce_film(timed):
        nonlocal reduce_ms
        if dist is None:
            return
        if timed:
            red0.record()
        dist.reduce(film, 0)
        if timed:
            red1.record(); red1.synchronize(); reduce_ms += red0.elapsed_time(red1)

    for _ in range(max(args.warmup, 3)):
        render_dev(film)
        reduce_film(False)
    barrier()
    scene.reset_stats()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ext_ms = ext_rays = ext_launches = shade_ms = shadow_ms = 0.0
    barrier()
    e0.record()
    for _ in range(args.steps):
        render_dev(film)
        reduce_film(True)
        st = scene.stats()
        ext_ms += st.extend_ms; ext_rays += st.extend_rays; ext_launches += st.extend_launches
        shade_ms += st.shade_ms; shadow_ms += st.shadow_ms
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    st = scene.stats()
    rays = float(st.closest_rays + st.shadow_rays)
    closest_rays, shadow_rays = float(st.closest_rays), float(st.shadow_rays)
    launches = int(st.kernel_launches)
    mean_radiance = float(torch.nanmean(film).item()) if rank == 0 else 0.0   # Whitted films hold the reference's NaN pixels
    if dist is not None:
        tt = torch.tensor([ms, reduce_ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms, reduce_ms = float(tt[0].item()), float(tt[1].item())
        rr = torch.tensor([rays, float(launches)], dtype=torch.float64, device="cuda"); dist.all_reduce(rr)
        rays, launches = float(rr[0].item()), int(rr[1].item())

    # T4: the N-GPU film (sum of the rank films, as just reduced onto rank 0) against the 1-GPU film of the same
    # spp, seed and stratification grid.  RNG keys depend on (pixel, global sample index) only, so every path is the same;
    # the films differ by float summation order.
    t4 = None
    if dist is not None and strong and rank == 0:
        sharded = film.clone()
        one = torch.zeros_like(film)
        if w["integrator"] == "bdpt":
            scene.render_bdpt_dev(cam, W.BdptParams(w["width"], w["height"], w["iterations"], 0, 10, 3, 1000, 0, 1, 0.0, 0), one.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
        else:
            fn = scene.render_whitted_dev if w["integrator"] == "whitted" else scene.render_pt_dev
            fn(cam, W.PtParams(w["width"], w["height"], w["spp"], w["depth"], 1000, 0, 1, 0.0), one.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        floor = 1e-3 * float(one.mean().item())
        rel = (sharded - one).abs() / torch.maximum(one.abs(), torch.tensor(floor, device="cuda"))
        t4 = {"max_rel_diff_vs_1gpu": float(rel.max().item()), "mean_rel_diff": float(rel.mean().item()),
              "bound": 1e-5, "ok": bool(rel.max().item() <= 1e-5),
              "note": "per-pixel |N-GPU - 1-GPU| / max(|1-GPU|, 1e-3 x mean radiance); same paths, float summation order differs"}
    if dist is not None:
        dist.barrier()

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

    # the dominant kernel timed ALONE: one more step with a single sub-pool, so that no launch of another stream overlaps the
    # extend launches whose CUDA-event durations are summed (with 2 sub-pools the sums exceed the wall time)
    alone = None
    if w["integrator"] in ("pt", "whitted") and world == 1:
        prev_sub = os.environ.get("WRT_SUBPOOLS")
        os.environ["WRT_SUBPOOLS"] = "1"
        try:
            render_dev(film); torch.cuda.synchronize()
            scene.reset_stats(); render_dev(film); torch.cuda.synchronize()
            sa = scene.stats()
            alone = {"extend_ms": sa.extend_ms, "extend_rays": float(sa.extend_rays), "launches": int(sa.extend_launches),
                     "render_ms": sa.last_render_ms}
        finally:
            if prev_sub is None:
                del os.environ["WRT_SUBPOOLS"]
            else:
                os.environ["WRT_SUBPOOLS"] = prev_sub

    value = rays / ms / 1e3
    total_samples = samples_per_step * args.steps * world
    unit_name = "iterations" if w["integrator"] == "bdpt" else "spp"
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "integrator": w["integrator"], "traversal": "pruned (bit-exact vs exact, tests/test_gpu_traversal.py)",
                   "per_gpu": ("full frame, %d %s per GPU, disjoint RNG streams; films summed by one NCCL reduce" if not strong else
                               "full frame, %d %s in total, dealt round-robin to the GPUs (same grid and seed as 1 GPU); films summed by one NCCL reduce") %
                              (w.get("spp", w.get("iterations")), unit_name),
                   "l2": "inputs larger than L2: the path pool (up to 2^26 slots x 176 B in concurrent sub-pools) is rewritten every bounce; the scene is meant to stay L2-resident",
                   "loop": "device-driven: every iteration reads its queue length from the counter bank its predecessor wrote; the host enqueues iterations "
                           "back to back and polls one batch late (pt_wavefront.cu)",
                   "prims": int(sc.n_prims), "scene_gen_s": round(scene_gen_s, 2), "kd_build_s": round(kd_build_s, 2), "upload_s": round(upload_s, 2)},
        "samples_per_s": total_samples / (ms / 1e3), "rays_per_sample": rays / total_samples,
        "mean_radiance": mean_radiance,
        "gpu_launches": launches,
        "clocks": clk,
        "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": unit,
                "h2d_bytes_per_step": int(C_sizeof_inputs(W, w)), "d2h_bytes_per_step": int(npix * 12),
                "ms_per_step": 1e3 * e2e_s / args.steps},
    }
    if world > 1:
        line["reduce_ms"] = reduce_ms / args.steps
        line["reduce"] = {"ms_per_step": reduce_ms / args.steps, "bytes": int(npix * 12), "what": "torch.distributed.reduce(SUM, fp32 film -> rank 0) over NCCL, "
                          "CUDA events around the collective, max over ranks (includes waiting for the slowest rank)"}
        line["t4_self_check"] = t4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    nm = ncu_metrics(args.workload) or {}
    if b_ray is not None and ext_ms > 0 and not sc_small_tree(scene, w):
        # dominant kernel = closest-hit traversal (k_pt_extend).  It is bound by ISSUE SLOTS and by the latency of dependent L1/L2
        # fetches, not by DRAM (the committed ncu capture: DRAM a few % of peak), so `bound` says so.  `achieved` / `frac` are the
        # kernel's OWN algorithmic bytes (what the PRUNED traversal fetches: 32 B per node visit, 48 B per leaf record, 40 B per ray)
        # over its launch durations, timed ALONE where possible; the reference-semantics figure SURVEY 8(d) defines is kept beside it.
        t_ms, t_rays, t_launch = (alone["extend_ms"], alone["extend_rays"], alone["launches"]) if alone else (ext_ms, ext_rays, ext_launches)
        own = (t_rays * b_ray_k) / (t_ms * 1e-3) / 1e9
        line["roofline"] = {
            "bound": "issue", "kernel": "k_pt_extend<pruned>",
            "achieved": own, "peak": peak, "unit": "GB/s", "frac": own / peak, "peak_source": peak_src,
            "what": "own-work algorithmic bytes (bytes_per_ray x rays) / summed CUDA-event durations of the extend launches of one step run with a "
                    "single sub-pool (no overlapping stream); against the measured HBM copy peak as the common yardstick — the binding limit is the "
                    "issue rate: see `issue`",
            "bytes_per_ray": b_ray_k, "visits_per_ray": visits_k,
            "avg_launch_ms": t_ms / max(t_launch, 1), "launches": int(t_launch), "rays_per_launch": t_rays / max(t_launch, 1),
            "extend_mrays_per_s": t_rays / t_ms / 1e3,
            "traffic": (nm.get("dram_bytes_per_ray") or 0) * t_rays / max(t_launch, 1) or None,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per ray of the committed ncu --set full capture (%s) x rays per launch" % nm.get("source", "none for this workload"),
            "issue": {"issue_active_pct": nm.get("issue_active_pct"), "threads_per_inst": nm.get("threads_per_inst"),
                      "useful_issue_frac": (nm["issue_active_pct"] / 100.0 * nm["threads_per_inst"] / 32.0) if nm.get("issue_active_pct") and nm.get("threads_per_inst") else None,
                      "warp_inst_per_ray": nm.get("warp_inst_per_ray"), "source": nm.get("source"),
                      "note": "issue slots in use x lanes doing useful work: the fraction of the SM's instruction throughput that advances rays"},
            "in_step": {"extend_ms_summed": ext_ms / args.steps, "kernel_share_of_step": ext_ms / ms,
                        "note": "summed launch durations inside the timed step; sub-pool streams overlap, so the share can exceed 1"},
            "reference_semantics": {"bytes_per_ray": b_ray, "visits_per_ray": visits, "achieved_GBps": (t_rays * b_ray) / (t_ms * 1e-3) / 1e9,
                                    "ratio_to_peak": (t_rays * b_ray) / (t_ms * 1e-3) / 1e9 / peak,
                                    "note": "SURVEY 8(d): bytes of the reference's FULL traversal (no early exit) of the same rays / this kernel's time; "
                                            "a ratio above 1 is work avoided by the bit-exact pruning, not bandwidth"}}
    elif w["integrator"] == "bdpt" and shade_ms > 0:
        # C4: 38 primitives — traversal is a handful of steps per ray; the dominant kernels are the per-vertex kernels (camera shade,
        # the (camera vertex, light vertex) connection kernel, light shade).  The committed ncu capture of the connection kernel says
        # what binds them: issue slots and divergence (two BSDF set-ups with IEEE divisions and one powf per pair), DRAM far from peak.
        # `achieved` keeps the HBM yardstick every line of this bench uses: algorithmic bytes of those kernels / their launch durations.
        per_step_closest, per_step_shadow = closest_rays / args.steps, shadow_rays / args.steps
        bytes_step = per_step_closest * (32 + 16 + 16 + 4 + 8 + 72) + per_step_shadow * (64 + 80 + 4 + 52 + 52)
        ach = bytes_step * args.steps / (shade_ms * 1e-3) / 1e9
        line["roofline"] = {
            "bound": "issue", "kernel": "k_bdpt_connect + k_bdpt_camera_shade + k_bdpt_light_shade", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "peak_source": peak_src,
            "what": "algorithmic bytes of the per-vertex kernels: per path vertex 76 B of state read + 72 B written; per connection 64 B light vertex + 80 B "
                    "camera record + 4 B pair entry read, 52 B queue entry written and 52 B read back by the occlusion kernel / summed CUDA-event durations "
                    "of those launches; against the measured HBM copy peak as the common yardstick — the binding limit is the issue rate: see `issue`",
            "avg_launch_ms": shade_ms / max(ext_launches, 1), "launches": int(ext_launches),
            "kernel_share_of_step": shade_ms / ms,
            "traffic": (nm.get("dram_bytes_per_launch") or None),
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture of one k_bdpt_connect launch (%s)" % nm.get("source", "none"),
            "issue": {"issue_active_pct": nm.get("issue_active_pct"), "threads_per_inst": nm.get("threads_per_inst"),
                      "useful_issue_frac": (nm["issue_active_pct"] / 100.0 * nm["threads_per_inst"] / 32.0) if nm.get("issue_active_pct") and nm.get("threads_per_inst") else None,
                      "source": nm.get("source")},
            "stage_ms_per_step": {"extend": ext_ms / args.steps, "shade+connect": shade_ms / args.steps, "occlusion+di": shadow_ms / args.steps},
            "traversal_own_work": {"visits_per_ray": visits_k, "bytes_per_ray": b_ray_k}}
    else:
        line["roofline"] = None
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(w, unit)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def sc_small_tree(scene, w):
    """Trees under 512 nodes (C4's Cornell box) run the plain per-thread traversal; their dominant kernel is the shade kernel."""
    return w["integrator"] == "bdpt"


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
    what = ("1 spp on every %d-th row of the %dx%d frame" % (rw["row_stride"], rw["width"], rw["height"])) if rw["integrator"] == "pt" else \
           ("1 %s of the whole %dx%d frame" % ("iteration" if rw["integrator"] == "bdpt" else "spp", rw["width"], rw["height"]))
    return {"value": rays / secs / 1e6, "unit": unit, "cores": 1, "kind": "reference",
            "sample": "%s (%d rays, %.1f s); reference KD build %.1f s not timed" % (what, rays, secs, build_s),
            "samples_per_s": samples / secs, "rays_per_sample": rays / max(samples, 1)}


if __name__ == "__main__":
    sys.exit(main())
