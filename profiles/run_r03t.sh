#!/bin/bash
# r03t: A/B of the score-only instantiation of the forward-noising kernel (no noise-matrix output compiled in)
mkdir -p gpurun_out
for v in base "" qsx5 qsx6 base ""; do
  if [ -z "$v" ]; then lib=""; tag=new; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "q_sample|error|Error" >> gpurun_out/r03t_probe.txt
done
cat gpurun_out/r03t_probe.txt
