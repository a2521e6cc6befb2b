WFS_LANES=${LANES:-4} timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('C1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
        for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('ms_per_step'), v.get('ms_phase_per_step'))
"
