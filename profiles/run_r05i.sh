#!/bin/bash
# r05i: auto evaluator, closed form first: series override inline (lib) vs out of line (autool)
mkdir -p gpurun_out
for v in "" autool "" autool; do
  if [ -z "$v" ]; then lib=""; tag=inline; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "score" >> gpurun_out/r05i_probe.txt
done
cut -c1-170 gpurun_out/r05i_probe.txt
