#!/bin/bash
# round 2, eighth GPU call: BDPT connect over a compacted pair list — parity, C4 line, launch list, ncu of the heaviest k_bdpt_connect launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tape.py tests/test_gpu_render.py -m gpu -q -x -k "bdpt or BDPT" > gpurun_out/pytest_gpu_r2h.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2h.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2h_c4.json 2>gpurun_out/bench_r2h_c4.err
python -c "
import json
j=json.loads(open('gpurun_out/bench_r2h_c4.json').read().strip().splitlines()[-1]); print('r2h c4: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']), j['roofline'].get('stage_ms_per_step',''))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c4_r2h.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4_launches_r2h.log 2>&1; echo "ncu c4 launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_bdpt_connect$ -s 9 -c 1 -f -o gpurun_out/prof_connect_r2h python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4_connect_r2h.log 2>&1; echo "ncu connect rc=$?"
