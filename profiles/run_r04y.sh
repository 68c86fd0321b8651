#!/bin/bash
# r04y: full ncu captures of the shipped two-row kernels at HEAD (forward noising, with the score, closed-form score, per-row-t reverse step)
T=r04y
mkdir -p gpurun_out
cap() { timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$1 -s 4 -c 1 -f -o gpurun_out/${T}_prof_$3 python tests/tools/probe_one.py $2 22 > gpurun_out/${T}_ncu_$3_stdout.log 2>&1; }
cap QSample2Op q_sample qsample2
cap QSample2Op q_sample_score qscore2
cap LogpScore2Op closed closed2
cap PStepRows2Op p_sample_rows pstep_rows2
ls -la gpurun_out | grep ${T}
