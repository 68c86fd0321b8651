#!/bin/bash
# r05a: shared-t reverse step per step index, HEAD vs the r04m library (uniform 1024-bucket shared-memory guide)
mkdir -p gpurun_out
for v in r04m "" r04m ""; do
  if [ -z "$v" ]; then lib=""; else lib=build/variants/libso3d_$v.so; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_pstep_t.py 24 >> gpurun_out/r05a_probe.txt
done
cat gpurun_out/r05a_probe.txt
