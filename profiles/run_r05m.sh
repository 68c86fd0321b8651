#!/bin/bash
# r05m: two-row forward noising: both lookups straight-line + one warp vote for the search (lib) vs a divergent region per row (novote)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "noising or q_sample or bench_size or guide" 2>&1 | tail -2
for v in novote "" novote ""; do
  if [ -z "$v" ]; then lib=""; tag=vote; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "\"q_sample" >> gpurun_out/r05m_probe.txt
done
cut -c1-170 gpurun_out/r05m_probe.txt
