#!/bin/bash
# round 2, seventh GPU call: BDPT connect kernel v2 (static partition, block-aggregated append): parity, C4 line, launch list, ncu of the
# heaviest k_bdpt_connect / k_bdpt_camera_shade launches; C3 / torus / c5_small with 8-byte stack entries as the default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tape.py tests/test_gpu_render.py -m gpu -q -x -k "bdpt or BDPT" > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2g.log
one() { # label workload steps env...
  label=$1; w=$2; steps=$3; shift 3
  env "$@" timeout 300 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_${label}_${w}.json 2>gpurun_out/bench_${label}_${w}.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_${label}_${w}.json').read().strip().splitlines()[-1]); print('$label $w: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']), j['roofline'].get('stage_ms_per_step',''))
except Exception as e: print('$label $w ERR', e)"
}
one r2g c4 3 X=1
one r2g c3 3 X=1
one r2g torus 3 X=1
one r2g c5_small 3 X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c4_r2g.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4_launches_r2g.log 2>&1; echo "ncu c4 launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_bdpt_connect$ -s 1 -c 1 -f -o gpurun_out/prof_connect_r2g python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4_connect.log 2>&1; echo "ncu connect rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_bdpt_camera_shade$ -s 1 -c 1 -f -o gpurun_out/prof_camshade_r2g python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4_camshade.log 2>&1; echo "ncu camshade rc=$?"
ls -la gpurun_out/*r2g*.ncu-rep
