#!/bin/bash
# r05j: full capture of the two-row auto score kernel (why 0.166 ms against 0.1505 for the closed form on eps <= 1 inputs?)
T=r05j
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:LogpScore2Op -s 4 -c 1 -f -o gpurun_out/${T}_prof_auto2 python tests/tools/probe_one.py auto 22 > gpurun_out/${T}_ncu_auto2_stdout.log 2>&1
ls -la gpurun_out | grep ${T}
