#!/bin/bash
mkdir -p gpurun_out
for v in se3bl se3c3 se3bl se3c3; do
  SO3D_LIB_PATH=build/variants/libso3d_$v.so timeout 300 python tests/tools/probe_engine.py 24 $v 2>&1 | grep -E "se3 q" >> gpurun_out/r04g_probe.txt
done
cut -c1-175 gpurun_out/r04g_probe.txt
