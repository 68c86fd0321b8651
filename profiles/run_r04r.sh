#!/bin/bash
# r04r: every map / backward / sampler op on the two-row warp-autonomous engine (TwoRow<Op>): full GPU tests, A/B vs one-row
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in one two one two; do
  if [ "$v" = one ]; then export SO3D_ROW_LANES=1 SO3D_LOGP_LANES=1; else unset SO3D_ROW_LANES SO3D_LOGP_LANES; fi
  timeout 300 python tests/tools/probe_engine.py 24 $v 2>&1 | grep -E "score|sample shared|log_|exp_|so3_|compose|rmat_dist" >> gpurun_out/r04r_probe.txt
done
cut -c1-175 gpurun_out/r04r_probe.txt
