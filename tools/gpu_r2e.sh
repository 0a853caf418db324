#!/bin/bash
# round 2, fifth GPU call: sub-pool staggering for frames without regeneration; coop threshold; BDPT camera_shade ncu
mkdir -p gpurun_out
one() { # label workload spp steps env...
  label=$1; w=$2; spp=$3; steps=$4; shift 4
  extra=""; [ "$spp" != "0" ] && extra="--spp $spp"
  env "$@" timeout 300 python bench.py --workload $w $extra --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_${label}_${w}_$spp.json 2>/dev/null
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_${label}_${w}_$spp.json').read().strip().splitlines()[-1]); print('$label $w spp=$spp: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']))
except Exception as e: print('$label $w ERR', e)"
}
for wts in "50,50" "65,35" "75,25" "85,15"; do
  one w$wts c3 8 5 WRT_SUBPOOL_WEIGHTS=$wts
  one w$wts c1 0 20 WRT_SUBPOOL_WEIGHTS=$wts
  one w$wts c3 16 5 WRT_SUBPOOL_WEIGHTS=$wts
done
one k3 c3 8 5 WRT_SUBPOOLS=3 WRT_SUBPOOL_WEIGHTS=50,30,20
one k3 c1 0 20 WRT_SUBPOOLS=3 WRT_SUBPOOL_WEIGHTS=50,30,20
one k4 c3 8 5 WRT_SUBPOOLS=4 WRT_SUBPOOL_WEIGHTS=40,30,20,10
one dflt c3 0 3 X=1
one dflt torus 0 3 X=1
one dflt c4 0 3 X=1
# BDPT: launch list + full capture of k_bdpt_camera_shade (largest launch: first camera iteration)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4_r2e.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4_launches.log 2>&1; echo "ncu c4 launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_bdpt_camera_shade$ -s 0 -c 1 -f -o gpurun_out/prof_camshade_r2e python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4_full.log 2>&1; echo "ncu c4 full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -2
