#!/bin/bash
# r01j: GPU tests at HEAD, bench line, launch list, full capture of the fused denoiser kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01j_pytest.log
tail -5 gpurun_out/r01j_pytest.log
timeout 600 python bench.py > gpurun_out/r01j_bench.json 2> gpurun_out/r01j_bench.err; tail -c 600 gpurun_out/r01j_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r01j_bench_reference.json 2>> gpurun_out/r01j_bench.err
timeout 300 python tests/tools/probe_denoiser.py 100000 16777216 > gpurun_out/r01j_probe_denoiser.log 2>&1; tail -8 gpurun_out/r01j_probe_denoiser.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:rotpredict_p_sample -s 4 -c 1 -f -o gpurun_out/prof_denoiser \
    python tests/tools/probe_denoiser.py 1000 4194304 > gpurun_out/ncu_denoiser_stdout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches_stdout.log 2>&1
ls -la gpurun_out
