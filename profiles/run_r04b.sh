#!/bin/bash
# r04b: full captures of the two-row forward-noising kernels
T=r04b
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:QSample2Op -s 4 -c 1 -f -o gpurun_out/${T}_prof_qsample2 \
    python tests/tools/probe_one.py q_sample 22 > gpurun_out/${T}_ncu_qsample_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:QSample2Op -s 4 -c 1 -f -o gpurun_out/${T}_prof_qscore2 \
    python tests/tools/probe_one.py q_sample_score 22 > gpurun_out/${T}_ncu_qscore_stdout.log 2>&1
ls -la gpurun_out | grep ${T}
