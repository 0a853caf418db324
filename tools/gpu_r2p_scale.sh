#!/bin/bash
# round 2, multi-GPU call: strong scaling of C3 / C5 (full, 10 M triangles + 100 k spheres) at N GPUs.  Usage: tools/gpu_r2p_scale.sh N "workloads"
N=${1:-4}; WL=${2:-"c3 c5"}
mkdir -p gpurun_out
for w in $WL; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --workload $w --steps 3 --warmup 3 > gpurun_out/bench_${w}_n${N}_r2p.json 2> gpurun_out/bench_${w}_n${N}_r2p.err; echo "bench $w n=$N rc=$?"
  python -c "
import json
try:
    j=json.loads([l for l in open('gpurun_out/bench_${w}_n${N}_r2p.json').read().strip().splitlines() if l.startswith('{')][-1]); print('$w n=$N: %.1f Mrays/s %.3f ms/step scaling=%s reduce_ms=%.3f t4_ok=%s kd %.1f s' % (j['value'], j['ms_per_step'], j['scaling'], j.get('reduce_ms', -1), (j.get('t4_self_check') or {}).get('ok'), j['config']['kd_build_s']))
except Exception as e: print('$w ERR', e)"
done
