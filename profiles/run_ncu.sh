#!/bin/bash
# Profiling recipe (run under gpurun, 1 GPU).  Outputs land in gpurun_out/; summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
# 1. launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches_stdout.log 2>&1
# 2. full capture of the dominant kernel (series logp+score) at a reduced n so the ~40 replays stay short
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:LogpScoreOp -s 3 -c 1 -f -o gpurun_out/prof_series \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-extra --n 4194304 > gpurun_out/ncu_series_stdout.log 2>&1
# 3. full capture of the HBM-bound fused kernels (reverse step, forward noising)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:PStepOp' -s 6 -c 1 -f -o gpurun_out/prof_steps \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_steps_stdout.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:QSampleOp' -s 6 -c 1 -f -o gpurun_out/prof_qsample \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_qsample_stdout.log 2>&1
# 4. the all-pairs MMD kernel (FP32-issue bound)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pair_sums_kernel' -s 3 -c 1 -f -o gpurun_out/prof_mmd \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_mmd_stdout.log 2>&1
ls -la gpurun_out
