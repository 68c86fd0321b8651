#!/bin/bash
# r03j: two-rows-per-thread shared-t reverse step (packed FP32): GPU tests + A/B against the one-row kernel
mkdir -p gpurun_out
T=r03j
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
for lanes in 1 2; do
  SO3D_PSTEP_LANES=$lanes timeout 300 python tests/tools/probe_engine.py 24 lanes$lanes 2>> gpurun_out/${T}.err | grep -E "p_sample shared|se3 p_sample shared|score auto" >> gpurun_out/${T}_probe.jsonl
done
cat gpurun_out/${T}_probe.jsonl | cut -c1-200
timeout 300 python tests/tools/probe_series.py 24 2>> gpurun_out/${T}.err | cut -c1-200
