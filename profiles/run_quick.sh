#!/bin/bash
# quick iteration: denoiser tests + probe timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_denoiser.py -m gpu -q > gpurun_out/quick_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/quick_pytest.log
tail -25 gpurun_out/quick_pytest.log
timeout 300 python tests/tools/probe_denoiser.py 100000 16777216 > gpurun_out/quick_probe.log 2>&1; tail -8 gpurun_out/quick_probe.log
