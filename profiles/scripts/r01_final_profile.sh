# Round-1 evidence run: plain bench first, then (same command) the launch list and one
# --set full capture per hot kernel.  One GPU.
set -x
CMD="python bench.py --steps 1 --warmup 1 --curves 262144 --no-e2e --no-cpu"
$CMD > gpurun_out/final_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu0.log 2>&1
$CMD > gpurun_out/final_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_pass -s 26 -c 1 -o gpurun_out/final_pass $CMD > gpurun_out/final_ncu1.log 2>&1
$CMD > gpurun_out/final_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_part2 -s 1 -c 1 -o gpurun_out/final_part2 $CMD > gpurun_out/final_ncu2.log 2>&1
$CMD > gpurun_out/final_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:splev_kernel -s 1 -c 1 -o gpurun_out/final_splev_fast $CMD > gpurun_out/final_ncu3.log 2>&1
$CMD > gpurun_out/final_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:shared_apply -s 1 -c 1 -o gpurun_out/final_shared $CMD > gpurun_out/final_ncu4.log 2>&1
ls -la gpurun_out/final_*
