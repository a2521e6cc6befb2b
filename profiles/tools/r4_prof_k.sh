# source-level profile of one kernel: KERNEL=regex TAG=name [SKIP] [COUNT]
F="--no-cpu-baseline --no-configs --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KERNEL" -s ${SKIP:-4} -c ${COUNT:-1} -o gpurun_out/${TAG}_k python bench.py --events 20000 $F --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu.log 2>&1; echo rc=$?
