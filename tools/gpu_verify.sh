#!/bin/bash
# Last check of a tree state: smoke(), the whole GPU suite, three bench lines.  Usage: tools/gpu_verify.sh <tag>
tag=${1:-v}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_$tag.json 2>/dev/null
timeout 300 python bench.py --workload whitted_torus --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_whitted_torus_$tag.json 2>/dev/null
timeout 300 python bench.py > gpurun_out/bench_c3_$tag.json 2>/dev/null
python - <<PY
import json
for f in ("bench_c3_$tag","bench_c4_$tag","bench_whitted_torus_$tag"):
    try:
        j=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, "%.1f %s %.1f ms/step e2e %.1f" % (j["value"], j["unit"], j["ms_per_step"], j["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
