# launch list (gpu__time_duration per launch) of a 20000-event C1 bench run
F="--no-cpu-baseline --no-configs --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG:-r4}_launches.csv python bench.py --events 20000 $F --steps 2 --warmup 1 > gpurun_out/${TAG:-r4}_launches.log 2>&1; echo rc=$?
