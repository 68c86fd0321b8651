#!/bin/bash
# r04q: closed-form / auto score on the two-row engine: parity, A/B vs the one-row kernel, resident-CTA caps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "closed_form_score or logp or score or host_score" 2>&1 | tail -3
for v in one "" lp6 lp10 lp12 one "" lp6 lp10 lp12; do
  unset SO3D_LOGP_LANES; lib=""; tag=two8
  if [ "$v" = one ]; then export SO3D_LOGP_LANES=1; tag=one; elif [ -n "$v" ]; then lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "\"score" >> gpurun_out/r04q_probe.txt
done
cut -c1-175 gpurun_out/r04q_probe.txt
