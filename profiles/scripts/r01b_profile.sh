# Round-1 (second session) evidence run on one B200: plain bench first, then (same command) the
# launch list, then one --set full capture per new/changed hot kernel, each after a plain run.
set -x
CMD="python bench.py --steps 1 --warmup 1 --curves 262144 --no-e2e --no-cpu"
$CMD > gpurun_out/r01b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01b_launches.csv $CMD > gpurun_out/r01b_ncu0.log 2>&1
S="python profiles/scripts/splev_only.py 4"
$S > gpurun_out/r01b_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:splev_fast -s 2 -c 1 -o gpurun_out/r01b_splev_fast $S > gpurun_out/r01b_ncu1.log 2>&1
G="python profiles/scripts/regrid_only.py 4096 1 0"
$G > gpurun_out/r01b_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:grid_plan -s 3 -c 1 -o gpurun_out/r01b_grid_plan $G > gpurun_out/r01b_ncu2.log 2>&1
$G > gpurun_out/r01b_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:grid_apply -s 6 -c 1 -o gpurun_out/r01b_grid_apply $G > gpurun_out/r01b_ncu3.log 2>&1
$G > gpurun_out/r01b_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:grid_resid -s 3 -c 1 -o gpurun_out/r01b_grid_resid $G > gpurun_out/r01b_ncu4.log 2>&1
$G > gpurun_out/r01b_plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:grid_bispev -s 0 -c 1 -o gpurun_out/r01b_grid_bispev $G > gpurun_out/r01b_ncu5.log 2>&1
ls -la gpurun_out/r01b_*
