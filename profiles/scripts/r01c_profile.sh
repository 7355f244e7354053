# Round-1 final evidence (after the regrid pipeline and shared-LSQ changes): launch list of the
# default-shaped bench + --set full captures of the kernels changed since r01b.  One B200.
set -x
CMD="python bench.py --steps 1 --warmup 1 --curves 262144 --no-e2e --no-cpu"
$CMD > gpurun_out/r01c_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/r01c_ncu0.log 2>&1
G="python profiles/scripts/regrid_only.py 4096 1 0"
$G > gpurun_out/r01c_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:grid_plan -s 3 -c 1 -o gpurun_out/r01c_grid_plan $G > gpurun_out/r01c_ncu1.log 2>&1
$G > gpurun_out/r01c_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:grid_apply -s 6 -c 1 -o gpurun_out/r01c_grid_apply $G > gpurun_out/r01c_ncu2.log 2>&1
S="python profiles/scripts/shared_only.py 2 0"
$S > gpurun_out/r01c_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:shared_apply2 -s 1 -c 1 -o gpurun_out/r01c_shared_exact $S > gpurun_out/r01c_ncu3.log 2>&1
S="python profiles/scripts/shared_only.py 2 1"
$S > gpurun_out/r01c_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:shared_apply2 -s 1 -c 1 -o gpurun_out/r01c_shared_fast $S > gpurun_out/r01c_ncu4.log 2>&1
ls -la gpurun_out/r01c_*
