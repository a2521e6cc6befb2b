# e2e leg of bench.py under different plain-row shares / expansion threads / core budgets
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=l['e2e']
print('dev ms', round(l['ms_per_step'],1), 'e2e ms', round(e['ms_per_step'],1), 'd2h GB', round(e['d2h_bytes_per_step']/1e9,2), 'share', e['plain_record_share'], 'ms_device', round(e['ms_device']))"; }
run WFS_PLAIN_FRACTION=0
run WFS_PLAIN_FRACTION=0.1
run WFS_PLAIN_FRACTION=0.2
run WFS_PLAIN_FRACTION=1
run WFS_EXPAND_THREADS=2 WFS_PLAIN_FRACTION=0
run WFS_EXPAND_THREADS=2 WFS_PLAIN_FRACTION=0.75
run WFS_EXPAND_THREADS=2 WFS_PLAIN_FRACTION=0.9
echo "== taskset 0-3"; 
taskset -c 0-3 env WFS_PLAIN_FRACTION=0 python bench.py --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=l['e2e']
print('dev ms', round(l['ms_per_step'],1), 'e2e ms', round(e['ms_per_step'],1), e['plain_record_share'])"
taskset -c 0-3 env WFS_PLAIN_FRACTION=0.8 python bench.py --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=l['e2e']
print('dev ms', round(l['ms_per_step'],1), 'e2e ms', round(e['ms_per_step'],1), e['plain_record_share'])"
taskset -c 0-3 python bench.py --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=l['e2e']
print('adaptive: dev ms', round(l['ms_per_step'],1), 'e2e ms', round(e['ms_per_step'],1), e['plain_record_share'])"
