#!/bin/bash
# r05g: per-row sampler with prefetch hooks: one-row (SO3D_ROW_LANES=1) vs two-row adaptor with hooks (default)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "sampler or sample or two_row or guide" 2>&1 | tail -2
for v in 1 0 1 0; do
  if [ "$v" = 1 ]; then export SO3D_ROW_LANES=1; else unset SO3D_ROW_LANES; fi
  timeout 300 python tests/tools/probe_secondary.py 24 lanes$v 2>&1 | grep "sample per-row" >> gpurun_out/r05g_probe.txt
done
cat gpurun_out/r05g_probe.txt
