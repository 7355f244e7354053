# r01f evidence (final state of this session: percur through the pass pipeline with register-resident
# two-block rotations): launch list of the default-shaped bench + --set full captures of the percur
# kernels.  One B200; every capture follows a plain run of the same command.
set -x
CMD="python bench.py --steps 1 --warmup 1 --curves 262144 --no-e2e --no-cpu"
$CMD > gpurun_out/r01f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01f_launches.csv $CMD > gpurun_out/r01f_ncu0.log 2>&1
Q="python profiles/scripts/curves_only.py percur 18 1"
$Q > gpurun_out/r01f_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01f_percur_launches.csv $Q > gpurun_out/r01f_ncu1.log 2>&1
$Q > gpurun_out/r01f_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:percur_pass -s 22 -c 1 -o gpurun_out/r01f_percur_pass $Q > gpurun_out/r01f_ncu2.log 2>&1
$Q > gpurun_out/r01f_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:percur_part2 -s 1 -c 1 -o gpurun_out/r01f_percur_part2 $Q > gpurun_out/r01f_ncu3.log 2>&1
cat gpurun_out/r01f_plain[1-3].log
ls -la gpurun_out/r01f_*
