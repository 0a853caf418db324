#!/bin/bash
# One GPU round-trip: parity tests, quick device probe, C3 + torus bench lines.  Usage: tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python tools/quick_perf.py --big > gpurun_out/perf_$tag.log 2>&1; echo "perf rc=$?"; cat gpurun_out/perf_$tag.log | grep -v "^\[wrt"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("gpurun_out/bench_c3_$tag.json",):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["value"], j["unit"], j["ms_per_step"], "e2e", j["e2e"]["value"], "roof", j["roofline"].get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 python bench.py --workload torus --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_torus_$tag.json 2> gpurun_out/bench_torus_$tag.err; echo "bench torus rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/bench_torus_$tag.json').read().strip().splitlines()[-1]); print('torus', j['value'], j['ms_per_step'])"
