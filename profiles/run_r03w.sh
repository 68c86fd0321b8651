#!/bin/bash
# r03w: GPU tests + bench at HEAD (float-format guide records) + capture of the two noising kernels
T=r03w
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:QSampleOp -s 4 -c 1 -f -o gpurun_out/${T}_prof_qscore \
    python tests/tools/probe_one.py q_sample_score 22 > gpurun_out/${T}_ncu_qscore_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:QSampleOp -s 4 -c 1 -f -o gpurun_out/${T}_prof_qsample \
    python tests/tools/probe_one.py q_sample 22 > gpurun_out/${T}_ncu_qsample_stdout.log 2>&1
ls -la gpurun_out | grep ${T}
