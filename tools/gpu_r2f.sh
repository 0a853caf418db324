#!/bin/bash
# round 2, sixth GPU call: BDPT camera shade split (parity + C4), 8-byte stack entries / streaming hints / L2 persistence A/B, coherence probe
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_render.py tests/test_gpu_tape.py -m gpu -q -x > gpurun_out/pytest_gpu_r2f.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2f.log
WRT_B200_LIB=libwrt_v_s8h.so timeout 900 python -m pytest tests/test_gpu_traversal.py -m gpu -x -q > gpurun_out/pytest_v_s8h.log 2>&1; echo "pytest(s8h) rc=$?"; tail -2 gpurun_out/pytest_v_s8h.log
one() { # label lib workload steps env...
  label=$1; lib=$2; w=$3; steps=$4; shift 4
  env "$@" WRT_B200_LIB=$lib timeout 300 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_${label}_${w}.json 2>gpurun_out/bench_${label}_${w}.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_${label}_${w}.json').read().strip().splitlines()[-1]); print('$label $w: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']))
except Exception as e: print('$label $w ERR', e)"
}
one r2f libwrt_b200.so c4 3 X=1
for v in b200 v_stack8 v_hints v_s8h; do
  for w in c3 torus c5_small; do one $v libwrt_$v.so $w 3 X=1; done
done
one l2n32 libwrt_b200.so c3 3 WRT_L2_PERSIST_MB=32
one l2n64 libwrt_b200.so c3 3 WRT_L2_PERSIST_MB=64
one l2r64 libwrt_b200.so c3 3 WRT_L2_PERSIST_MB=64 WRT_L2_PERSIST_WHAT=recs
one l2r64s8h libwrt_v_s8h.so c3 3 WRT_L2_PERSIST_MB=64 WRT_L2_PERSIST_WHAT=recs
timeout 600 python tools/coherence_probe.py > gpurun_out/coherence_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/coherence_probe.log | tail -8
