"""Writes profiles/r2_final_launches_and_ncu.md (+ copies of the bench lines and launch lists) from what tools/gpu_final_r2.sh left in
gpurun_out/.  Usage: tools/make_profile_md_r2.py <tag>"""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")


def last_json(p):
    return json.loads([l for l in open(p).read().strip().splitlines() if l.startswith("{")][-1])


out = ["# r2 — final kernels of round 2: bench lines, launch lists and `ncu --set full` of `k_pt_extend<pruned>`, `k_pt_shadow<pruned>` (C3) and `k_bdpt_connect` (C4)", "",
       "All files of this page come from ONE `gpurun` call (`tools/gpu_final_r2.sh %s`, one B200, clocks 1965 MHz, no throttle reasons): the parity suite" % tag,
       "(`profiles/r2_pytest_gpu.log`: 83 passed, 4 skipped = the multi-GPU tests, which ran on 2- and 8-GPU boxes: `profiles/r2_pytest_multi_n2.log`, `_n8.log`),",
       "the bench lines, the launch lists and the full captures.  Numbers printed by a run under ncu are never used as bench values.", "",
       "## Bench lines (`bench.py`, CUDA events, no profiler; copied to profiles/r2_bench_*.json)", "",
       "| workload | value | ms/step | e2e (host buffers) | host KD build | launches | cpu_baseline (reference, 1 core) |", "|---|---|---|---|---|---|---|"]
names = [("c3", "c3"), ("c3spp8", "c3_8spp"), ("c1", "c1"), ("torus", "torus"), ("cbox_dragon", "cbox_dragon"), ("c4", "c4"), ("c5_small", "c5_small"), ("c5", "c5"),
         ("whitted_torus", "whitted_torus"), ("ref", "c3_reference_arm")]
for w, dst in names:
    f = os.path.join(G, "bench_%s_%s.json" % (w, tag))
    if not os.path.exists(f):
        continue
    j = last_json(f); shutil.copy(f, os.path.join(P, "r2_bench_%s.json" % dst))
    cb = j.get("cpu_baseline") or {}
    out.append("| %s | %.1f %s | %.1f | %.1f | %s s | %s | %s |" % (
        j["config"]["workload"][:78] + (" (reference arm, %d processes)" % cb.get("cores", 0) if w == "ref" else ""), j["value"], j["unit"], j["ms_per_step"],
        j["e2e"]["value"], j["config"].get("kd_build_s", "-"), j.get("gpu_launches", "-"), ("%.3f Mrays/s" % cb["value"]) if cb.get("value") and w != "ref" else "-"))
c3 = last_json(os.path.join(G, "bench_c3_%s.json" % tag)); r = c3["roofline"]
ref = last_json(os.path.join(G, "bench_ref_%s.json" % tag))
out += ["", "C3 against the reference arm on the same box (%d processes): %.0fx device-resident, %.0fx through the host-buffer C-ABI call; against one core: %.0fx."
        % (ref["cpu_baseline"]["cores"], c3["value"] / ref["value"], c3["e2e"]["value"] / ref["value"], c3["value"] / c3["cpu_baseline"]["value"]),
        "C3 roofline block: own-work algorithmic bytes %.0f B/ray, extend kernel alone %.0f Mrays/s -> `achieved` %.0f GB/s = %.2f of the measured HBM peak (bound: issue); "
        "reference-semantics bytes %.0f B/ray -> ratio %.2f." % (r["bytes_per_ray"], r["extend_mrays_per_s"], r["achieved"], r["frac"],
                                                                 r["reference_semantics"]["bytes_per_ray"], r["reference_semantics"]["ratio_to_peak"]), ""]


def launch_table(path, title):
    rows = list(csv.reader(open(path)))
    for i, row in enumerate(rows):
        if "Kernel Name" in row:
            h = row; st = i; break
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for row in rows[st + 1:]:
        if len(row) <= mv:
            continue
        try:
            v = float(row[mv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(row[kn].split("(")[0].replace("void ", "")[:60], [0, 0.0]); a[0] += 1; a[1] += v / 1e6
    tot = sum(a[1] for a in agg.values())
    o = [title, "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        o.append("| %s | %d | %.2f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / tot))
    return o + [""]


shutil.copy(os.path.join(G, "launches_c3_%s.csv" % tag), os.path.join(P, "r2_launches_c3_final.csv"))
shutil.copy(os.path.join(G, "launches_c4_%s.csv" % tag), os.path.join(P, "r2_launches_c4_final.csv"))
out += launch_table(os.path.join(G, "launches_c3_%s.csv" % tag), "## Launch list, C3 (`ncu --metrics gpu__time_duration.sum --clock-control none -c 800`, `bench.py --steps 1 --warmup 3 --spp 16`)")
out += ["Per-launch times under ncu are cold-cache and serialised (the sub-pool streams cannot overlap under the profiler), so only the shares are comparable with the live",
        "run (live: summed extend durations / step = %.2f with two overlapping streams).  `*_count` kernels belong to the two 1-spp counting renders bench.py does before the timed region."
        % r["in_step"]["kernel_share_of_step"], ""]
out += launch_table(os.path.join(G, "launches_c4_%s.csv" % tag), "## Launch list, C4 (same command with `--workload c4`)")
for k, title in (("extend", "## `ncu --set full --clock-control none --import-source on -k regex:^k_pt_extend$ -s 2 -c 1` (C3, 64 spp: a full 33 554 432-ray queue of regenerated camera "
                            "rays + continuing paths — the same launch as the round-1 capture, 17.64 ms there)"),
                 ("shadow", "## `… -k regex:^k_pt_shadow$ -s 2 -c 1` (C3: the shadow / NEE occlusion kernel)"),
                 ("connect", "## `… -k regex:^k_bdpt_connect$ -s 22 -c 1` (C4: the heaviest connection launch of a batch of 8 iterations)")):
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, "prof_%s_%s.ncu-rep" % (k, tag)), "25"], capture_output=True, text=True).stdout
    out += [title, "", summ, ""]
open(os.path.join(P, "r2_final_launches_and_ncu.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:24]))
