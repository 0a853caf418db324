"""Writes profiles/<round>_final_launches_and_ncu.md (+ copies of the bench lines and the launch list) from the files
tools/gpu_final.sh left in gpurun_out/.  Usage: tools/make_profile_md.py <tag> <round-prefix>   e.g.  r1d r1"""
import collections, csv, json, os, shutil, subprocess, sys
tag, rnd = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")

def last_json(path):
    for l in reversed(open(path).read().strip().splitlines()):
        if l.startswith("{"):
            return json.loads(l)
    raise RuntimeError("no JSON line in " + path)

out = ["# %s — final kernels: bench lines, launch list and `ncu --set full` of `k_pt_extend<pruned>` (scheduler 3, default build)" % rnd, ""]
out += ["All files of this page come from ONE `gpurun` call (`tools/gpu_final.sh %s`, one B200): the parity suite, the bench lines, the" % tag,
        "launch list and the full capture.  Numbers printed by a run under ncu are never used as bench values.", ""]
out += ["## Bench lines (`bench.py`, CUDA events, no profiler; copied to profiles/%s_bench_*.json)" % rnd, "",
        "| workload | value | ms/step | e2e (host buffers) | clocks MHz | launches |", "|---|---|---|---|---|---|"]
for w in ("c3", "torus", "cbox_dragon", "c4", "c5_small", "c5", "ref"):
    f = os.path.join(G, "bench_%s_%s.json" % (w, tag))
    if not os.path.exists(f):
        continue
    j = last_json(f)
    shutil.copy(f, os.path.join(P, "%s_bench_%s.json" % (rnd, w if w != "ref" else "c3_reference_arm")))
    clk = j.get("clocks") or {}
    out.append("| %s | %.1f %s | %.1f | %.1f | %s | %s |" % (j["config"]["workload"][:70] + (" (reference arm, %d processes)" % j["cpu_baseline"]["cores"] if w == "ref" else ""),
               j["value"], j["unit"], j["ms_per_step"], j["e2e"]["value"], clk.get("sm_mhz"), j.get("gpu_launches", "-")))
c3 = last_json(os.path.join(G, "bench_c3_%s.json" % tag))
r = c3["roofline"]
out += ["", "C3 roofline block: reference-semantics bytes %.0f B/ray -> `achieved` %.0f GB/s (frac %.2f of the measured HBM peak: more than 1 because PRUNED"
        % (r["bytes_per_ray"], r["achieved"], r["frac"]),
        "traversal skips most of that work); the kernels' own algorithmic bytes %.0f B/ray -> %.0f GB/s; cpu_baseline (reference, 1 core): %.3f Mrays/s."
        % (r["kernel_own_work"]["bytes_per_ray"], r["kernel_own_work"]["achieved_GBps"], c3["cpu_baseline"]["value"]),
        "Summed extend-launch durations / step wall time = %.2f (two sub-pool streams overlap)." % r["kernel_share_of_step"], ""]

rows = list(csv.reader(open(os.path.join(G, "launches_%s.csv" % tag))))
shutil.copy(os.path.join(G, "launches_%s.csv" % tag), os.path.join(P, "%s_launches_final.csv" % rnd))
for i, row in enumerate(rows):
    if "Kernel Name" in row:
        h = row; start = i; break
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for row in rows[start + 1:]:
    if len(row) <= mv: continue
    try: v = float(row[mv].replace(",", ""))
    except ValueError: continue
    a = agg.setdefault(row[kn].split("(")[0][:60], [0, 0.0]); a[0] += 1; a[1] += v / 1e6
tot = sum(a[1] for a in agg.values())
out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -c 800`, `bench.py --steps 1 --warmup 3 --spp 16`)", "",
        "Per-launch times under ncu are cold-cache and serialised (the sub-pool streams cannot overlap under the profiler), so only the shares",
        "are comparable with the live run.  `*_count` kernels belong to the two 1-spp counting renders bench.py does before the timed region.", "",
        "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| %s | %d | %.2f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / tot))
ext = sum(a[1] for k, a in agg.items() if "k_pt_extend<" in k and "count" not in k)
timed = sum(a[1] for k, a in agg.items() if "count" not in k and "at::" not in k)
out += ["", "Share of `k_pt_extend<1>` among the render kernels (extend, shade, shadow, init): %.0f %% under ncu; live run: extend durations / step = %.2f with"
        % (100 * ext / timed, r["kernel_share_of_step"]), "overlap, i.e. the same kernel dominates, followed by `k_pt_shadow<1>`.", ""]

summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, "prof_extend_%s.ncu-rep" % tag), "25"],
                      capture_output=True, text=True).stdout
out += ["## `ncu --set full --clock-control none --import-source on -k regex:^k_pt_extend$ -s 2 -c 1` (`bench.py --steps 1 --warmup 3`, C3, 64 spp:",
        "the third extend launch = second iteration of sub-pool 0 = a full 33 554 432-ray queue of regenerated camera rays + continuing paths)", "", summ]
open(os.path.join(P, "%s_final_launches_and_ncu.md" % rnd), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
