#!/bin/bash
# round 2, fourth GPU call: cooperative tail + shadow stream — parity, then A/B on C1 / C3@8spp / C3 / torus / c5_small
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_render.py tests/test_gpu_tape.py -m gpu -q -x > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2d.log
run() { # name lib env...
  name=$1; lib=$2; shift 2
  for spec in "c1 0 20" "c3 8 5" "c3 0 3" "torus 0 3" "c5_small 0 3"; do
    set -- $spec "$@"; w=$1; spp=$2; steps=$3; shift 3
    extra=""; [ "$spp" != "0" ] && extra="--spp $spp"
    env "$@" WRT_B200_LIB=$lib timeout 300 python bench.py --workload $w $extra --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_${name}_${w}_$spp.json 2>/dev/null
    python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_${name}_${w}_$spp.json').read().strip().splitlines()[-1]); print('$name $w spp=$spp: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']))
except Exception as e: print('$name $w ERR', e)"
  done
}
run default libwrt_b200.so X=1
run nocoop libwrt_v_nocoop.so X=1
run coop8 libwrt_v_coop8.so X=1
run noshadowstream libwrt_b200.so WRT_SHADOW_STREAM=0
python tools/c1_as_shipped.py > gpurun_out/c1_as_shipped_r2d.log 2>&1; cat gpurun_out/c1_as_shipped_r2d.log
