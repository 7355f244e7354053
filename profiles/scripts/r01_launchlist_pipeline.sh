# launch list of the pass pipeline at two batch sizes (time-only metric: one replay per kernel)
for C in 131072 1048576; do
CMD="python bench.py --steps 1 --warmup 1 --curves $C --no-e2e --no-cpu --no-splev"
$CMD > gpurun_out/plain_$C.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pipe_$C.csv -k regex:'curfit|p2_' $CMD > gpurun_out/ncu_pipe_$C.log 2>&1
done
ls -la gpurun_out/*.csv
