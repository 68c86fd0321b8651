#!/bin/bash
# r04s: two output stages for TwoRow<Op> (lib) vs one (os1, which also gives the closed-form score two stages); one-row for reference
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "two_row" 2>&1 | tail -3
for v in one two os1 one two os1; do
  unset SO3D_ROW_LANES SO3D_LOGP_LANES; lib=""
  if [ "$v" = one ]; then export SO3D_ROW_LANES=1 SO3D_LOGP_LANES=1; fi
  if [ "$v" = os1 ]; then lib=build/variants/libso3d_os1.so; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $v 2>&1 | grep -E "score|sample shared|log_|exp_|so3_|compose|rmat_dist" >> gpurun_out/r04s_probe.txt
done
cut -c1-175 gpurun_out/r04s_probe.txt
