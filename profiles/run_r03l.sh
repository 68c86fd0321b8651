#!/bin/bash
mkdir -p gpurun_out
T=r03l
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -12 gpurun_out/${T}_pytest.log
SO3D_QSAMPLE_LANES=1 SO3D_PSTEP_LANES=1 timeout 300 python tests/tools/probe_engine.py 24 lanes1 2>> gpurun_out/${T}.err | grep -E "q_sample|p_sample shared" >> gpurun_out/${T}_probe.jsonl
for lib in shipped build/variants/libso3d_qs4.so build/variants/libso3d_qs6.so; do
  if [ "$lib" = shipped ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$lib; fi
  timeout 300 python tests/tools/probe_engine.py 24 $(basename $lib .so) 2>> gpurun_out/${T}.err | grep -E "q_sample|p_sample shared" >> gpurun_out/${T}_probe.jsonl
done
cat gpurun_out/${T}_probe.jsonl | cut -c1-175
