#!/bin/bash
# r05f: SE(3) per-row-t reverse step with prefetch hooks: one-row (SO3D_ROW_LANES=1) vs two-row adaptor with hooks (default)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "se3 or two_row" 2>&1 | tail -2
for v in 1 0 1 0; do
  if [ "$v" = 1 ]; then export SO3D_ROW_LANES=1; else unset SO3D_ROW_LANES; fi
  timeout 300 python tests/tools/probe_secondary.py 24 lanes$v 2>&1 | grep "se3" >> gpurun_out/r05f_probe.txt
done
cat gpurun_out/r05f_probe.txt
