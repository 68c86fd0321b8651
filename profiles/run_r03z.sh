#!/bin/bash
# r03z: per-row-t reverse step through the prefetch hooks; resident-CTA cap 3 (lib) / 4 / 5; base = HEAD~ (no hooks, cap 3)
mkdir -p gpurun_out
for v in base "" ps4 ps5 base "" ps4 ps5; do
  if [ -z "$v" ]; then lib=""; tag=ps3; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "p_sample per-row" >> gpurun_out/r03z_probe.txt
done
cut -c1-175 gpurun_out/r03z_probe.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "p_sample or reverse or p_step or pstep" 2>&1 | tail -3
