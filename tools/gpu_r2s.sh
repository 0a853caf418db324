#!/bin/bash
# round 2: whole-warp traversal of the last rays (par_traverse): off / 2 (default) / 4 / 8 rays
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_traversal.py -m gpu -x -q 2>&1 | tail -2
for lib in libwrt_v_par0.so libwrt_b200.so libwrt_v_par4.so libwrt_v_par8.so; do
  echo "== $lib"
  WRT_B200_LIB=$lib python tools/tail_probe.py 2>&1 | tail -4
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
