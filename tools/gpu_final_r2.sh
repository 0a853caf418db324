#!/bin/bash
# Final evidence set of round 2 (one B200): parity suite, bench lines of every workload (both arms for C3), launch lists, ncu --set full of
# k_pt_extend / k_pt_shadow (C3) and k_bdpt_connect (C4).   Usage: tools/gpu_final_r2.sh <tag>
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 900 python bench.py > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; echo "bench c3 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "bench ref rc=$?"
for w in c4 c5; do
  timeout 1500 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err; echo "bench $w rc=$?"
done
for w in c1 torus cbox_dragon c5_small whitted_torus; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err; echo "bench $w rc=$?"
done
timeout 300 python bench.py --workload c3 --spp 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3spp8_$tag.json 2>/dev/null
python tools/c1_as_shipped.py > gpurun_out/c1_as_shipped_$tag.log 2>&1; cat gpurun_out/c1_as_shipped_$tag.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_$tag.json")):
    try:
        j=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1]); print(f, "%.2f %s, %.1f ms/step, e2e %.2f, kd %s s" % (j["value"], j["unit"], j["ms_per_step"], j["e2e"]["value"], j["config"].get("kd_build_s")))
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c3_$tag.csv python bench.py --steps 1 --warmup 3 --spp 16 --no-cpu-baseline > gpurun_out/ncu_launches_c3_$tag.log 2>&1; echo "ncu c3 launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c4_$tag.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_c4_$tag.log 2>&1; echo "ncu c4 launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_pt_extend$ -s 2 -c 1 -f -o gpurun_out/prof_extend_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_extend_$tag.log 2>&1; echo "ncu extend rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_pt_shadow$ -s 2 -c 1 -f -o gpurun_out/prof_shadow_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_shadow_$tag.log 2>&1; echo "ncu shadow rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_bdpt_connect$ -s 22 -c 1 -f -o gpurun_out/prof_connect_$tag python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_connect_$tag.log 2>&1; echo "ncu connect rc=$?"
ls -la gpurun_out/*_$tag.ncu-rep
