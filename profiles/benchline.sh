timeout 200 python bench.py --steps 20 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), {k[:20]:round(v['ms_per_launch']*1000) for k,v in d['kernels'].items() if v['ms_per_launch']>0.03})"
