#!/bin/bash
# r04d: is it the fifth resident CTA that slows the branch-free two-row noising kernel? (SO3D_CTAS_PER_SM caps the persistent grid)
mkdir -p gpurun_out
for c in 3 4 5 3 4 5; do
  SO3D_CTAS_PER_SM=$c timeout 300 python tests/tools/probe_engine.py 24 two_c$c 2>&1 | grep -E "q_sample (per|\+)" >> gpurun_out/r04d_probe.txt
  SO3D_CTAS_PER_SM=$c SO3D_LIB_PATH=build/variants/libso3d_w2br.so timeout 300 python tests/tools/probe_engine.py 24 w2br_c$c 2>&1 | grep -E "q_sample (per|\+)" >> gpurun_out/r04d_probe.txt
done
cut -c1-175 gpurun_out/r04d_probe.txt
