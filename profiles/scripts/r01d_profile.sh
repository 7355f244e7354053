# r01d evidence (parametric + periodic curves added; fit pipeline templated on idim): launch list of
# the default-shaped bench + --set full captures of the fit pipeline kernels (idim=1 and idim=2),
# percur_kernel and curev_kernel.  One B200; every capture follows a plain run of the same command.
set -x
CMD="python bench.py --steps 1 --warmup 1 --curves 262144 --no-e2e --no-cpu"
$CMD > gpurun_out/r01d_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01d_launches.csv $CMD > gpurun_out/r01d_ncu0.log 2>&1
F="python profiles/scripts/curves_only.py fit 20 1"
$F > gpurun_out/r01d_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01d_fit_launches.csv $F > gpurun_out/r01d_ncu1.log 2>&1
$F > gpurun_out/r01d_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_pass -s 30 -c 1 -o gpurun_out/r01d_fit_pass $F > gpurun_out/r01d_ncu2.log 2>&1
$F > gpurun_out/r01d_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_part2 -s 1 -c 1 -o gpurun_out/r01d_fit_part2 $F > gpurun_out/r01d_ncu3.log 2>&1
P="python profiles/scripts/curves_only.py parcur 18 1"
$P > gpurun_out/r01d_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_pass -s 30 -c 1 -o gpurun_out/r01d_parcur_pass $P > gpurun_out/r01d_ncu4.log 2>&1
Q="python profiles/scripts/curves_only.py percur 16 1"
$Q > gpurun_out/r01d_plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:percur_kernel -s 1 -c 1 -o gpurun_out/r01d_percur $Q > gpurun_out/r01d_ncu5.log 2>&1
C="python profiles/scripts/curves_only.py curev 18 1"
$C > gpurun_out/r01d_plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curev_kernel -s 1 -c 1 -o gpurun_out/r01d_curev $C > gpurun_out/r01d_ncu6.log 2>&1
cat gpurun_out/r01d_plain[1-6].log
ls -la gpurun_out/r01d_*
