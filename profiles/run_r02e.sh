#!/bin/bash
# r02e: device-seed forward noising + graphed training step -- tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_seed or training_step or q_sample" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
tail -30 gpurun_out/r02e_pytest.log
timeout 600 python tests/tools/probe_train.py --graph --cpu > gpurun_out/r02e_probe_train.jsonl 2> gpurun_out/r02e.err; cat gpurun_out/r02e_probe_train.jsonl | cut -c1-400; tail -c 500 gpurun_out/r02e.err
