#!/bin/bash
# Final evidence set of a round: parity suite, bench lines (both arms), ncu launch list + one --set full capture.
# Usage: tools/gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "bench ref rc=$?"
for w in torus cbox_dragon c4 c5_small c5; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err; echo "bench $w rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_$tag.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.2f %s, %.1f ms/step, e2e %.2f" % (j["value"], j["unit"], j["ms_per_step"], j["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --spp 16 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
bash tools/gpu_ncu.sh $tag
