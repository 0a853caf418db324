#!/bin/bash
# round 2: tail-loop knobs after the lane-group passes: tail entry at <= 4 (default) / 8 / 16 rays, 12 / 24 (default) / 48 node steps per tail round
mkdir -p gpurun_out
for lib in libwrt_b200.so libwrt_v_coop8.so libwrt_v_coop16.so libwrt_v_ts12.so libwrt_v_ts48.so; do
  for spec in "c1 0 20" "c3 8 5" "c3 0 3" "torus 0 3" "c5_small 0 3"; do
    set -- $spec; w=$1; spp=$2; steps=$3
    extra=""; [ "$spp" != "0" ] && extra="--spp $spp"
    WRT_B200_LIB=$lib timeout 300 python bench.py --workload $w $extra --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_${lib}_${w}_$spp.json 2>/dev/null
    python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_${lib}_${w}_$spp.json').read().strip().splitlines()[-1]); print('$lib $w spp=$spp: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']))
except Exception as e: print('$lib $w ERR', e)"
  done
done
