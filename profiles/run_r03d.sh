#!/bin/bash
# r03d: sanitizer re-checks after the fixes (racecheck / synccheck / memcheck over every kernel family) + ncu captures of the
# kernels that changed this round (series with guard, reverse step, forward noising, one-launch reverse process)
mkdir -p gpurun_out
T=r03d
for tool in racecheck synccheck memcheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 10 python tests/tools/sanitize_target.py 1 257 1300 > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done|denoiser n|Barrier error|Race reported" gpurun_out/${T}_sanitizer_$tool.log | sort | uniq -c | head -8
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sweep --no-eager --no-accuracy > gpurun_out/${T}_ncu_launches_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:LogpScoreOp -s 3 -c 1 -f -o gpurun_out/${T}_prof_series \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-extra --no-accuracy --n 4194304 > gpurun_out/${T}_ncu_series_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:PStepOp -s 4 -c 1 -f -o gpurun_out/${T}_prof_pstep \
    python tests/tools/probe_one.py p_sample 22 > gpurun_out/${T}_ncu_pstep_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:QSampleOp -s 4 -c 1 -f -o gpurun_out/${T}_prof_qsample \
    python tests/tools/probe_one.py q_sample 22 > gpurun_out/${T}_ncu_qsample_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:p_sample_loop_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_loop \
    python tests/tools/probe_one.py loop 21 > gpurun_out/${T}_ncu_loop_stdout.log 2>&1
ls -la gpurun_out | grep ${T}
