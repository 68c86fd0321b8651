#!/bin/bash
# r04l: full GPU tests at HEAD (per-warp input slices default) + register-cap variants of the two-row kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
for v in "" se3w2 qs2c3 "" se3w2 qs2c3; do
  if [ -z "$v" ]; then lib=""; tag=head; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "\"(q_sample|p_sample per|se3 q)" >> gpurun_out/r04l_probe.txt
done
cut -c1-175 gpurun_out/r04l_probe.txt
