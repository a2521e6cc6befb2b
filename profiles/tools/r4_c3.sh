for L in 4 5; do WFS_LANES=$L timeout 400 python bench.py --config C3 --events 40000 --no-cpu-baseline --no-e2e --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lanes $L', d['value'], d['ms_per_step'], d.get('ms_phase_per_step'))
"; done
