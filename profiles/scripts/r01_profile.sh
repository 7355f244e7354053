python -m pytest tests/test_gpu_parity.py -q -m gpu -k "shared" > gpurun_out/pt.log 2>&1; tail -3 gpurun_out/pt.log
CMD="python bench.py --steps 1 --warmup 1 --curves 131072 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_kernel -s 1 -c 1 -o gpurun_out/prof_fit $CMD > gpurun_out/ncu2.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:splev_kernel -s 1 -c 1 -o gpurun_out/prof_splev_fast $CMD > gpurun_out/ncu3.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:splev_kernel -s 3 -c 1 -o gpurun_out/prof_splev_exact $CMD > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-600; ls -la gpurun_out/
