#!/bin/bash
# r02j: the whole reverse process in ONE launch (loop mode of the fused denoiser kernel) -- tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_denoiser.py -m gpu -q -x > gpurun_out/r02j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest.log
tail -25 gpurun_out/r02j_pytest.log
timeout 300 python tests/tools/probe_denoiser.py 1000 16777216 2>&1 | tail -4
timeout 600 python tests/tools/probe_loop.py > gpurun_out/r02j_probe_loop.jsonl 2> gpurun_out/r02j.err; cat gpurun_out/r02j_probe_loop.jsonl; tail -c 500 gpurun_out/r02j.err
