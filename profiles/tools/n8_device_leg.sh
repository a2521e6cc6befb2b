# device-resident leg at 8 ranks: what makes a rank slow?  (per-rank ms in the JSON line)
run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', '%.3g' % l['value'], 'ms', round(l['ms_per_step'],1), 'per rank', l.get('ms_per_step_per_rank'), 'host phase', l['ms_phase_per_step']['host_scheduler_truth'])"; PORT=$((PORT+1)); }
PORT=29530
run WFS_NOOP=1
run WFS_BLOCKING_SYNC=0
run WFS_BENCH_NO_CLOCKS=1
run WFS_BENCH_NO_CLOCKS=1 WFS_BLOCKING_SYNC=0
run WFS_BENCH_NO_CLOCKS=1 WFS_LANES=2
