#!/bin/bash
# A/B of library variants on the GPU box: tools/gpu_variants.sh <check_variant> <variant> [<variant> ...]
# (variants = names given to tools/build_variant.sh -o libwrt_v_<name>.so; "default" = libwrt_b200.so)
chk=$1; shift
mkdir -p gpurun_out
if [ "$chk" != "none" ]; then
  WRT_B200_LIB=libwrt_v_$chk.so timeout 900 python -m pytest tests/test_gpu_traversal.py -m gpu -x -q > gpurun_out/pytest_v_$chk.log 2>&1; echo "pytest($chk) rc=$?"; tail -2 gpurun_out/pytest_v_$chk.log
fi
for v in "$@"; do
  lib=libwrt_v_$v.so; [ "$v" = "default" ] && lib=libwrt_b200.so
  for w in ${WL:-c3 torus}; do
    WRT_B200_LIB=$lib timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v_${v}_$w.json 2> gpurun_out/bench_v_${v}_$w.err
    python -c "
import json,sys
try:
    j=json.loads(open('gpurun_out/bench_v_${v}_$w.json').read().strip().splitlines()[-1]); print('$v $w %.1f Mrays/s %.1f ms/step' % (j['value'], j['ms_per_step']))
except Exception as e: print('$v $w ERR', e)"
  done
done
