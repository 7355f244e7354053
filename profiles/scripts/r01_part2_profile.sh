timeout 600 python -m pytest tests -m gpu -q -x -k "batch or mncurf or golden or full_size" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 2 --warmup 3 --no-cpu --no-splev --no-e2e > gpurun_out/bench6.log 2>&1; tail -c 1200 gpurun_out/bench6.log | cut -c1-400
CMD="python bench.py --steps 1 --warmup 1 --curves 262144 --no-e2e --no-cpu --no-splev"
$CMD > gpurun_out/plainp2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_part2 -s 1 -c 1 -o gpurun_out/prof_part2 $CMD > gpurun_out/ncu_p2.log 2>&1
$CMD > gpurun_out/plainp1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_pass -s 26 -c 1 -o gpurun_out/prof_pass $CMD > gpurun_out/ncu_p1.log 2>&1
ls -la gpurun_out/*.ncu-rep
