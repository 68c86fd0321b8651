#!/bin/bash
mkdir -p gpurun_out
T=r03o
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_launch or shard_invariance or p_sample or small_batch or series" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
for lib in shipped build/variants/libso3d_is1c7.so build/variants/libso3d_is1c6.so build/variants/libso3d_is1os2.so; do
  if [ "$lib" = shipped ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$lib; fi
  timeout 300 python tests/tools/probe_engine.py 24 $(basename $lib .so) 2>> gpurun_out/${T}.err | grep -E "p_sample shared t" | grep -v se3 >> gpurun_out/${T}_probe.jsonl
  timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_launch" 2>&1 | tail -1
done
unset SO3D_LIB_PATH
cat gpurun_out/${T}_probe.jsonl | cut -c1-170
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:series_warp python tests/tools/probe_one.py series_small 12 2>/dev/null | grep series_warp | tail -3
