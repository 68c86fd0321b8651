#!/bin/bash
# r03u: full capture of the score-only forward-noising kernel (QSampleOp<1,0,0>) and of the plain one
T=r03u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:QSampleOp -s 4 -c 1 -f -o gpurun_out/${T}_prof_qscore \
    python tests/tools/probe_one.py q_sample_score 22 > gpurun_out/${T}_ncu_qscore_stdout.log 2>&1
ls -la gpurun_out | grep ${T}
