# source-level profile of k_group_analyse (the 4 size-class launches of one device batch)
F="--no-cpu-baseline --no-configs --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_group_analyse.*" -s ${SKIP:-16} -c ${COUNT:-4} -o gpurun_out/${TAG:-r4b}_analyse python bench.py --events 20000 $F --steps 1 --warmup 1 > gpurun_out/${TAG:-r4b}_ncu.log 2>&1; echo rc=$?
