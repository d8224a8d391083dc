timeout 300 python bench.py --steps 40 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['value'])
tot=0
for k,v in d['kernels'].items():
    c=v['launches_per_step']*v['ms_per_launch']*1000; tot+=c
    print(f\"{k[:44]:44s} {v['launches_per_step']:5.2f} x {v['ms_per_launch']*1000:7.1f} = {c:7.1f} us/step\")
print('sum', round(tot,1))"
