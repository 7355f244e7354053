CMD="python bench.py --steps 1 --warmup 1 --curves 131072 --no-e2e --no-cpu"
$CMD > gpurun_out/plains.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:splev_kernel -s 1 -c 1 -o gpurun_out/prof_splev_fast2 $CMD > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/prof_splev_fast2.ncu-rep
