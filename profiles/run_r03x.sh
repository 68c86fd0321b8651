#!/bin/bash
# r03x: per-instruction captures of the shared-t reverse step (two-row kernel), the one-launch loop and the auto score kernel
T=r03x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:PStep2Op -s 4 -c 1 -f -o gpurun_out/${T}_prof_pstep2 \
    python tests/tools/probe_one.py p_sample 22 > gpurun_out/${T}_ncu_pstep_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:p_sample_loop_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_loop \
    python tests/tools/probe_one.py loop 21 > gpurun_out/${T}_ncu_loop_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:LogpScoreOp -s 4 -c 1 -f -o gpurun_out/${T}_prof_auto \
    python tests/tools/probe_one.py auto 22 > gpurun_out/${T}_ncu_auto_stdout.log 2>&1
ls -la gpurun_out | grep ${T}
