#!/bin/bash
# round 2, third GPU call: TOS-cache A/B, sub-pool count for small frames, C1 launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_render.py -m gpu -q -x > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2c.log
WL="c3 torus c5_small" bash tools/gpu_variants.sh none default notos
for k in 1 2 4 8; do
  WRT_SUBPOOLS=$k timeout 300 python bench.py --workload c3 --spp 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3spp8_sub$k.json 2>/dev/null
  python -c "
import json; j=json.loads(open('gpurun_out/bench_c3spp8_sub$k.json').read().strip().splitlines()[-1]); print('c3 spp8 subpools=$k: %.1f Mrays/s %.2f ms/step' % (j['value'], j['ms_per_step']))"
  WRT_SUBPOOLS=$k timeout 300 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c1_sub$k.json 2>/dev/null
  python -c "
import json; j=json.loads(open('gpurun_out/bench_c1_sub$k.json').read().strip().splitlines()[-1]); print('c1 subpools=$k: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']))"
done
python tools/c1_as_shipped.py > gpurun_out/c1_as_shipped_r2c.log 2>&1; cat gpurun_out/c1_as_shipped_r2c.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c1_r2c.csv python bench.py --workload c1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
tail -40 gpurun_out/launches_c1_r2c.csv | cut -d, -f5,12- | tail -40
