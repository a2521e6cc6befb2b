# round-2 (session 4) check of a changed fused back end: A/B and replay tests, short bench; optional fuzz
set -x
F="--no-cpu-baseline --no-configs --no-e2e"
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_replay.py tests/test_gpu_deterministic.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python bench.py $F --steps 3 --warmup 2 > gpurun_out/r4_bench.log 2>&1; echo rc=$?
python - <<'P'
import json
for l in open('gpurun_out/r4_bench.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('ms_per_step', d['ms_per_step'], 'value', d['value'], d['roofline']['kernels'], d['ms_phase_per_step'])
P
if [ -n "$FUZZ" ]; then timeout 600 python profiles/tools/fuzz_fused.py $FUZZ 2>&1 | tail -5; fi
