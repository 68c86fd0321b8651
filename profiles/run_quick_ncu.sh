#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:rotpredict_p_sample -s 4 -c 1 -f -o gpurun_out/prof_denoiser2 \
    python tests/tools/probe_denoiser.py 1000 4194304 > gpurun_out/ncu_denoiser_stdout.log 2>&1
tail -3 gpurun_out/ncu_denoiser_stdout.log
