# r01g evidence (end of this session): --set full captures of the kernels changed after r01d --
# fit part 2 with the fused g := a copy, the idim=2 pipeline at 5 CTAs/SM, the shared-memory curev kernel.
# One B200; every capture follows a plain run of the same command.
set -x
F="python profiles/scripts/curves_only.py fit 20 1"
$F > gpurun_out/r01g_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_part2 -s 1 -c 1 -o gpurun_out/r01g_fit_part2 $F > gpurun_out/r01g_ncu1.log 2>&1
$F > gpurun_out/r01g_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_pass -s 25 -c 1 -o gpurun_out/r01g_fit_pass $F > gpurun_out/r01g_ncu2.log 2>&1
P="python profiles/scripts/curves_only.py parcur 19 1"
$P > gpurun_out/r01g_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_pass -s 28 -c 1 -o gpurun_out/r01g_parcur_pass $P > gpurun_out/r01g_ncu3.log 2>&1
$P > gpurun_out/r01g_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curfit_part2 -s 1 -c 1 -o gpurun_out/r01g_parcur_part2 $P > gpurun_out/r01g_ncu4.log 2>&1
C="python profiles/scripts/curves_only.py curev 19 1"
$C > gpurun_out/r01g_plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:curev_kernel -s 1 -c 1 -o gpurun_out/r01g_curev $C > gpurun_out/r01g_ncu5.log 2>&1
cat gpurun_out/r01g_plain[1-5].log
ls -la gpurun_out/r01g_*.ncu-rep
