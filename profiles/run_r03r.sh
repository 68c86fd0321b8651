#!/bin/bash
mkdir -p gpurun_out
T=r03r
timeout 600 python -m pytest tests/test_gpu_se3.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
SO3D_SE3_PSTEP_LANES=1 timeout 300 python tests/tools/probe_engine.py 24 lanes1 2>> gpurun_out/${T}.err | grep -E "se3" >> gpurun_out/${T}_probe.jsonl
for lib in shipped build/variants/libso3d_pf0.so build/variants/libso3d_pf1c4.so build/variants/libso3d_pf1c2.so; do
  if [ "$lib" = shipped ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$lib; fi
  timeout 300 python tests/tools/probe_engine.py 24 $(basename $lib .so) 2>> gpurun_out/${T}.err | grep -E "se3" >> gpurun_out/${T}_probe.jsonl
done
cat gpurun_out/${T}_probe.jsonl | cut -c1-175
