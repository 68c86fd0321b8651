#!/bin/bash
# r04n: float-format shared-memory guide (shared-row kernels): GPU tests, then A/B vs HEAD (base)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in base "" base ""; do
  if [ -z "$v" ]; then lib=""; tag=new; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "shared" >> gpurun_out/r04n_probe.txt
done
cut -c1-175 gpurun_out/r04n_probe.txt
SO3D_LIB_PATH=build/variants/libso3d_base.so timeout 300 python tests/tools/probe_loop.py 2>&1 | tail -3
timeout 300 python tests/tools/probe_loop.py 2>&1 | tail -3
