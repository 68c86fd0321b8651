#!/bin/bash
# r05d: the secondary per-row kernels on the one-row engine (SO3D_ROW_LANES=1) vs the TwoRow adaptor (=2)
mkdir -p gpurun_out
for v in 1 2 1 2; do
  SO3D_ROW_LANES=$v timeout 300 python tests/tools/probe_secondary.py 24 lanes$v >> gpurun_out/r05d_probe.txt
done
cat gpurun_out/r05d_probe.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "se3 or sampler or two_row" 2>&1 | tail -2
