#!/bin/bash
mkdir -p gpurun_out
T=r03q
timeout 600 python -m pytest tests/test_gpu_denoiser.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 200 python tests/tools/probe_denoiser_time.py 24 | tee -a gpurun_out/${T}_denoiser.jsonl
SO3D_LIB_PATH=build/variants/libso3d_issuer.so timeout 200 python tests/tools/probe_denoiser_time.py 24 | tee -a gpurun_out/${T}_denoiser.jsonl
