# r02 evidence: launch lists (fit pipeline alone; default-shaped bench) and --set full captures of the two fit
# kernels, exact splev and curev.  One B200; every capture follows a plain run of the same command.
set -x
F="python profiles/scripts/curves_only.py fit 20 1"
$F > gpurun_out/r02_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fit_launches.csv $F > gpurun_out/r02_ncu1.log 2>&1
$F > gpurun_out/r02_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_pass -s 4 -c 1 -o gpurun_out/r02_fit_pass $F > gpurun_out/r02_ncu2.log 2>&1
$F > gpurun_out/r02_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_part2 -s 0 -c 1 -o gpurun_out/r02_fit_part2 $F > gpurun_out/r02_ncu3.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --curves 262144 --no-e2e --no-cpu"
$CMD > gpurun_out/r02_plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu4.log 2>&1
S="python profiles/scripts/splev_only.py 3 0"
$S > gpurun_out/r02_plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:splev_kernel -s 1 -c 1 -o gpurun_out/r02_splev_exact $S > gpurun_out/r02_ncu5.log 2>&1
C="python profiles/scripts/curves_only.py curev 18 1"
$C > gpurun_out/r02_plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curev_kernel -s 1 -c 1 -o gpurun_out/r02_curev $C > gpurun_out/r02_ncu6.log 2>&1
cat gpurun_out/r02_plain[1-6].log | tail -20
ls -la gpurun_out/r02_*
