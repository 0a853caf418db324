#!/bin/bash
# round 2, second GPU call: device-driven wavefront — full parity suite, C3 / C4 / torus bench lines, C1 as shipped
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/pytest_gpu_$tag.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_$tag.log | head -20
python tools/c1_as_shipped.py > gpurun_out/c1_as_shipped_$tag.log 2>&1; cat gpurun_out/c1_as_shipped_$tag.log
for w in c3 c4 torus; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err; echo "bench $w rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_$tag.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.2f %s, %.1f ms/step, e2e %.2f, launches %d" % (j["value"], j["unit"], j["ms_per_step"], j["e2e"]["value"], j["gpu_launches"]))
    except Exception as e: print(f, "ERR", e)
PY
