#!/bin/bash
# round 2, first GPU call: full parity suite on the new tree + a fresh C3 bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2a.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_r2a.json 2> gpurun_out/bench_c3_r2a.err; echo "bench c3 rc=$?"
python - <<PY
import json
for f in ["gpurun_out/bench_c3_r2a.json"]:
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.2f %s, %.1f ms/step, e2e %.2f" % (j["value"], j["unit"], j["ms_per_step"], j["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
