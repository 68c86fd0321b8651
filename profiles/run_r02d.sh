#!/bin/bash
# r02d: reverse loop as one CUDA graph (device-resident Philox seed) -- tests + loop timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_denoiser.py tests/test_cabi_symbols.py -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log
tail -30 gpurun_out/r02d_pytest.log
timeout 600 python tests/tools/probe_loop.py > gpurun_out/r02d_probe_loop.jsonl 2> gpurun_out/r02d.err; cat gpurun_out/r02d_probe_loop.jsonl; tail -c 800 gpurun_out/r02d.err
