#!/bin/bash
# round 2, ninth GPU call: ray suspension — full parity suite, then tail-budget A/B on C1 / C3@8spp / C3 / torus / c5_small
# (historical: WRT_TAIL_BUDGET drove the ray-suspension experiment, whose code was removed again — commit 1f02113, profiles/r2_experiments.md)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x -s > gpurun_out/pytest_gpu_r2i.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|handed over" gpurun_out/pytest_gpu_r2i.log | tail -4
one() { # label workload spp steps env...
  label=$1; w=$2; spp=$3; steps=$4; shift 4
  extra=""; [ "$spp" != "0" ] && extra="--spp $spp"
  env "$@" timeout 300 python bench.py --workload $w $extra --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_${label}_${w}_$spp.json 2>gpurun_out/bench_${label}_${w}_$spp.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_${label}_${w}_$spp.json').read().strip().splitlines()[-1]); print('$label $w spp=$spp: %.1f Mrays/s %.3f ms/step launches %d' % (j['value'], j['ms_per_step'], j['gpu_launches']))
except Exception as e: print('$label $w ERR', e)"
}
for b in 0 16 48 128; do
  one tb$b c1 0 20 WRT_TAIL_BUDGET=$b
  one tb$b c3 8 5 WRT_TAIL_BUDGET=$b
  one tb$b c3 0 3 WRT_TAIL_BUDGET=$b
done
for b in 0 48; do
  one tb$b torus 0 3 WRT_TAIL_BUDGET=$b
  one tb$b c5_small 0 3 WRT_TAIL_BUDGET=$b
done
WRT_TAIL_BUDGET=48 python tools/c1_as_shipped.py > gpurun_out/c1_as_shipped_r2i.log 2>&1; cat gpurun_out/c1_as_shipped_r2i.log
