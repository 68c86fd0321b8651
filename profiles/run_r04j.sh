#!/bin/bash
# r04j: per-warp single-stage input (kWarpIn) for the two-row noising / per-row-t reverse step kernels vs the two-stage ring
mkdir -p gpurun_out
SO3D_LIB_PATH=build/variants/libso3d_warpin.so timeout 900 python -m pytest tests -m gpu -q -x -k "two_row or q_sample or p_sample" 2>&1 | tail -3
for v in ring warpin ring warpin; do
  if [ "$v" = warpin ]; then lib=build/variants/libso3d_warpin.so; else lib=""; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $v 2>&1 | grep -E "\"(q_sample|p_sample per)" >> gpurun_out/r04j_probe.txt
done
cut -c1-175 gpurun_out/r04j_probe.txt
