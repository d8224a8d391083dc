"""Pretty-print a bench.py JSON line: python profiles/show.py <file>"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches'] if k in d})
print('e2e', (d.get('e2e') or {}).get('value'), 'clocks', d.get('clocks'))
r = d.get('roofline') or {}
print('roofline', {k: r.get(k) for k in ('kernel', 'achieved', 'frac')}, 'group', r.get('fitness_replication_group'))
tot = 0
for k, v in (d.get('kernels') or {}).items():
    print(f"{k:32s} {v['launches_per_step']:4.1f} {v['ms_per_launch']*1000:8.1f} us  {v.get('GBps','')}")
    tot += v['launches_per_step'] * v['ms_per_launch']
print('sum of kernels per step (ms):', round(tot, 4))
if 'cpu_baseline' in d: print('cpu', d['cpu_baseline'])
