#!/bin/bash
# r04u: auto's series branch out of line (new) vs inline (base = HEAD): tests, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "score or logp or q_sample or noising or smoke or series" 2>&1 | tail -3
for v in base "" base ""; do
  if [ -z "$v" ]; then lib=""; tag=new; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "score" >> gpurun_out/r04u_probe.txt
done
cut -c1-175 gpurun_out/r04u_probe.txt
