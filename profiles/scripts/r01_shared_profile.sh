CMD="python bench.py --steps 1 --warmup 1 --curves 131072 --no-e2e --no-cpu --no-splev"
$CMD > gpurun_out/plainsh.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:shared_apply -s 1 -c 1 -o gpurun_out/prof_shared $CMD > gpurun_out/ncu_sh.log 2>&1
ls -la gpurun_out/prof_shared.ncu-rep
