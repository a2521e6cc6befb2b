# e2e leg of bench.py: lanes x expansion threads
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --no-configs --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=l['e2e']
print('dev ms', round(l['ms_per_step'],1), 'e2e ms', round(e['ms_per_step'],1), 'ms_device', round(e['ms_device']))"; }
run WFS_LANES=4
run WFS_LANES=3
run WFS_LANES=2
run WFS_LANES=3 WFS_EXPAND_THREADS=12
run WFS_LANES=4 WFS_EXPAND_THREADS=12
run WFS_LANES=3 WFS_BLOCKING_SYNC=1
