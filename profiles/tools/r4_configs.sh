# launch lists of the C2 and C4 bench legs (bounded samples)
for C in C2 C4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches_$C.csv python bench.py --config $C --no-cpu-baseline --no-e2e --steps 1 --warmup 1 ${EV:+--events $EV} > gpurun_out/r4_launches_$C.log 2>&1; echo rc=$?
tail -c 400 gpurun_out/r4_launches_$C.log
done
