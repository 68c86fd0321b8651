#!/bin/bash
# r04a: forward noising on two rows per thread, warp-autonomous (rowwise_kernel_w2): parity test, then A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "two_row_noising or q_sample or noise" 2>&1 | tail -5
for v in one "" qs2c5 one "" qs2c5; do
  if [ "$v" = one ]; then lib=""; tag=one; export SO3D_QS_LANES=1; elif [ -z "$v" ]; then lib=""; tag=two; unset SO3D_QS_LANES; else lib=build/variants/libso3d_$v.so; tag=$v; unset SO3D_QS_LANES; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "q_sample" >> gpurun_out/r04a_probe.txt
done
cut -c1-175 gpurun_out/r04a_probe.txt
