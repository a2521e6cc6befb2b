# round-2 (session 4) check of a changed fused back end: A/B and replay tests, short bench per variant; optional fuzz
# VARIANTS="A=1 B=2;A=0" runs the bench once per ';'-separated environment
F="--no-cpu-baseline --no-configs ${BENCH_FLAGS---no-e2e}"
if [ -z "$NOTEST" ]; then
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_replay.py tests/test_gpu_deterministic.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -15
fi
IFS=';' read -ra VS <<< "${VARIANTS:-default=1}"
i=0
for v in "${VS[@]}"; do
  i=$((i+1))
  env $v timeout 300 python bench.py $F --steps 3 --warmup 2 > gpurun_out/r4_bench_$i.log 2>&1; echo "variant [$v] rc=$?"
  python - gpurun_out/r4_bench_$i.log <<'P'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        k = d['roofline']['kernels']
        print('  ms_per_step %.2f value %.3e e2e %s | alone: analyse %.2f records %.2f | phases %s' % (
            d['ms_per_step'], d['value'], ('%.1f ms' % d['e2e']['ms_per_step']) if 'e2e' in d and d['e2e'] else '-',
            k['k_group_analyse']['ms_per_step'], k['k_group_records']['ms_per_step'], d['ms_phase_per_step']))
P
done
if [ -n "$FUZZ" ]; then timeout 900 python profiles/tools/fuzz_fused.py $FUZZ 2>&1 | tail -5; fi
