#!/bin/bash
# r03v: float-format guide records (2051 per row) + elect.sync store issue + score-only noising kernel: tests, then A/B vs HEAD~ (base)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r03v_pytest.log 2>&1; tail -5 gpurun_out/r03v_pytest.log
for v in base "" base ""; do
  if [ -z "$v" ]; then lib=""; tag=new; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "sample|p_s|error|Error" >> gpurun_out/r03v_probe.txt
done
cat gpurun_out/r03v_probe.txt | cut -c1-150
