#!/bin/bash
# r04h: shared-t reverse step on the warp-autonomous two-row engine (SO3D_PSTEP_ENGINE=w2) vs the CTA-synchronous two-row kernel
mkdir -p gpurun_out
SO3D_PSTEP_ENGINE=w2 timeout 900 python -m pytest tests -m gpu -q -x -k "p_sample or reverse or loop" 2>&1 | tail -3
for v in cta2 w2 cta2 w2; do
  if [ "$v" = w2 ]; then export SO3D_PSTEP_ENGINE=w2; else unset SO3D_PSTEP_ENGINE; fi
  timeout 300 python tests/tools/probe_engine.py 24 $v 2>&1 | grep -E "p_sample shared" >> gpurun_out/r04h_probe.txt
  for c in 4 5; do SO3D_CTAS_PER_SM=$c timeout 300 python tests/tools/probe_engine.py 24 ${v}_c$c 2>&1 | grep -E "^.{0,40}p_sample shared" >> gpurun_out/r04h_probe.txt; done
done
cut -c1-175 gpurun_out/r04h_probe.txt
