#!/bin/bash
# r04e: two-row warp-autonomous kernels for forward noising (trait: branch-free with the score) and the per-row-t reverse step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in one two one two; do
  if [ "$v" = one ]; then export SO3D_QS_LANES=1 SO3D_PS_LANES=1; else unset SO3D_QS_LANES SO3D_PS_LANES; fi
  timeout 300 python tests/tools/probe_engine.py 24 $v 2>&1 | grep -E "sample" >> gpurun_out/r04e_probe.txt
done
cut -c1-175 gpurun_out/r04e_probe.txt
