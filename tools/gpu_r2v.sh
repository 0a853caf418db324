#!/bin/bash
# round 2, last call: extend walks continuing paths first (variant) against camera rays first (default); then the final sanity of the committed tree:
# parity suite, smoke(), default bench line
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libwrt_v_contfirst.so libwrt_b200.so; do
  for spec in "c3 0 3" "torus 0 3" "c5_small 0 3"; do
    set -- $spec; w=$1; spp=$2; steps=$3
    WRT_B200_LIB=$lib timeout 300 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_${lib}_${w}_$spp.json 2>/dev/null
    python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_${lib}_${w}_$spp.json').read().strip().splitlines()[-1]); print('$lib $w: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']))
except Exception as e: print('$lib $w ERR', e)"
  done
done
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2v.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r2v.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_c3_r2v.json 2>gpurun_out/bench_c3_r2v.err; echo "bench rc=$?"; python -c "
import json
j=json.loads(open('gpurun_out/bench_c3_r2v.json').read().strip().splitlines()[-1]); print('default bench: %.1f Mrays/s, e2e %.1f, frac %.2f, cpu %.3f' % (j['value'], j['e2e']['value'], j['roofline']['frac'], j['cpu_baseline']['value']))"
