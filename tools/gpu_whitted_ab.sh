one() { label=$1; shift; env "$@" timeout 300 python bench.py --workload whitted_torus --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print('$label: %.1f Mrays/s %.2f ms/step launches %d' % (j['value'], j['ms_per_step'], j['gpu_launches']))"; }
one default X=1
one noshadowstream WRT_SHADOW_STREAM=0
one sub1 WRT_SUBPOOLS=1
one sub4 WRT_SUBPOOLS=4
