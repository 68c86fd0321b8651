#!/bin/bash
# r04k: SE(3) noising two-row with per-warp input slices (4 and 3 resident CTAs) vs the one-row kernel
mkdir -p gpurun_out
for v in one two two3 one two two3; do
  lib=""; unset SO3D_SE3_QS_LANES
  if [ "$v" = two ]; then export SO3D_SE3_QS_LANES=2; fi
  if [ "$v" = two3 ]; then export SO3D_SE3_QS_LANES=2; lib=build/variants/libso3d_se3w3.so; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $v 2>&1 | grep -E "\"se3 q" >> gpurun_out/r04k_probe.txt
done
cut -c1-175 gpurun_out/r04k_probe.txt
