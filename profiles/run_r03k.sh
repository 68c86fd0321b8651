#!/bin/bash
mkdir -p gpurun_out
T=r03k
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_launch or shard_invariance or p_sample" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
SO3D_PSTEP_LANES=1 timeout 300 python tests/tools/probe_engine.py 24 lanes1 2>> gpurun_out/${T}.err | grep -E "p_sample shared" >> gpurun_out/${T}_probe.jsonl
for lib in shipped build/variants/libso3d_os1.so build/variants/libso3d_t256.so; do
  if [ "$lib" = shipped ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$lib; fi
  timeout 300 python tests/tools/probe_engine.py 24 $(basename $lib .so) 2>> gpurun_out/${T}.err | grep -E "p_sample shared" >> gpurun_out/${T}_probe.jsonl
done
cat gpurun_out/${T}_probe.jsonl | cut -c1-170
