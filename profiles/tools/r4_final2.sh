# default bench line and reference arm of the final build
( time timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r4_bench_default.json 2> gpurun_out/r4_bench_default.err ) 2>&1 | tail -3; tail -c 300 gpurun_out/r4_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r4_bench_reference.json 2> gpurun_out/r4_bench_reference.err ) 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches.csv python bench.py --events 20000 --no-cpu-baseline --no-configs --no-e2e --steps 2 --warmup 1 > gpurun_out/r4_launches.log 2>&1
