set -x
F="--no-cpu-baseline --no-configs --no-e2e"
timeout 300 python bench.py --events 20000 $F --steps 2 --warmup 1 > gpurun_out/r4a_b0.log 2>&1; echo rc=$?
tail -c 600 gpurun_out/r4a_b0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_group_analyse -s 8 -c 4 -o gpurun_out/r4a_analyse python bench.py --events 20000 $F --steps 1 --warmup 1 > gpurun_out/r4a_ncu.log 2>&1; echo rc=$?
tail -3 gpurun_out/r4a_ncu.log
for L in 5 6; do WFS_LANES=$L timeout 300 python bench.py $F --steps 3 --warmup 2 > gpurun_out/r4a_lanes$L.log 2>&1; tail -c 300 gpurun_out/r4a_lanes$L.log | grep -o '"ms_per_step": [0-9.]*'; done
timeout 300 python bench.py $F --steps 3 --warmup 2 > gpurun_out/r4a_lanes4.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4a_lanes4.log | head -1
