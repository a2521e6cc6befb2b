# final measurements of the session: whole GPU suite, smoke, fuzz, ncu --set full of the fused back end (-> traffic JSON
# that bench.py reads), default bench line, reference arm, launch list
set -x
F="--no-cpu-baseline --no-configs --no-e2e"
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python profiles/tools/fuzz_fused.py 23 400 2>&1 | tail -1
timeout 400 python profiles/tools/fuzz_replay.py 23 300 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_group_(analyse.*|records)" -s 10 -c 13 -o gpurun_out/r4_fused python bench.py --events 20000 $F --steps 1 --warmup 1 > gpurun_out/r4_fused_ncu.log 2>&1
ncu -i gpurun_out/r4_fused.ncu-rep --page raw --csv > gpurun_out/r4_fused_raw.csv 2>/dev/null
python profiles/tools/fused_traffic.py profiles/r4_fused_kernels_ncu.txt profiles/r2_fused_ncu.json < gpurun_out/r4_fused_raw.csv
cp profiles/r2_fused_ncu.json gpurun_out/r2_fused_ncu.json
python profiles/ncu_metrics.py < gpurun_out/r4_fused_raw.csv > gpurun_out/r4_fused_kernels_ncu.txt
( time timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r4_bench_default.json 2> gpurun_out/r4_bench_default.err ) 2>&1 | tail -3; tail -c 300 gpurun_out/r4_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r4_bench_reference.json 2> gpurun_out/r4_bench_reference.err ) 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches.csv python bench.py --events 20000 $F --steps 2 --warmup 1 > gpurun_out/r4_launches.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_photons|k_instr_truth" -s 4 -c 2 -o gpurun_out/r4_front python bench.py --events 20000 $F --steps 1 --warmup 1 > gpurun_out/r4_front_ncu.log 2>&1
ncu -i gpurun_out/r4_front.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_metrics.py > gpurun_out/r4_front_kernels_ncu.txt
ls -la gpurun_out | head -40
