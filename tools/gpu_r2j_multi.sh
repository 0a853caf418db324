#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): the library's own multi-GPU path (wrt_init) and the torchrun strong-scaling bench with T4 self-check
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -q -s -k "multi_gpu or all_gpus" > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"; grep -E "passed|failed|GPUs:" gpurun_out/pytest_multi_n$N.log | tail -6
for w in c3 c5_small c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --workload $w --steps 3 --warmup 3 > gpurun_out/bench_${w}_n${N}_r2j.json 2> gpurun_out/bench_${w}_n${N}_r2j.err; echo "bench $w n=$N rc=$?"
  python -c "
import json
try:
    j=json.loads([l for l in open('gpurun_out/bench_${w}_n${N}_r2j.json').read().strip().splitlines() if l.startswith('{')][-1]); print('$w n=$N: %.1f Mrays/s %.3f ms/step scaling=%s reduce_ms=%.3f t4=%s' % (j['value'], j['ms_per_step'], j['scaling'], j.get('reduce_ms', -1), j.get('t4_self_check')))
except Exception as e: print('$w ERR', e)"
done
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_n1_r2j.json 2>/dev/null
python -c "
import json
j=json.loads(open('gpurun_out/bench_c3_n1_r2j.json').read().strip().splitlines()[-1]); print('c3 n=1 same box: %.1f Mrays/s %.3f ms/step' % (j['value'], j['ms_per_step']))"
