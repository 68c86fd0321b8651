#!/bin/bash
# r04f: SE(3) noising on the two-row warp-autonomous engine: parity + A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "two_row or se3" 2>&1 | tail -3
for v in one two one two; do
  if [ "$v" = one ]; then export SO3D_SE3_QS_LANES=1; else unset SO3D_SE3_QS_LANES; fi
  timeout 300 python tests/tools/probe_engine.py 24 $v 2>&1 | grep -E "se3|p_sample shared" >> gpurun_out/r04f_probe.txt
done
cut -c1-175 gpurun_out/r04f_probe.txt
