#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r01w_probe_engine.jsonl
run() { timeout 300 python tests/tools/probe_engine.py 24 $1 2>> gpurun_out/r01w.err | grep "p_sample" >> gpurun_out/r01w_probe_engine.jsonl; }
run keyed
SO3D_LIB_PATH=$PWD/build/variants/libso3d_plain.so run plain
run keyed
SO3D_LIB_PATH=$PWD/build/variants/libso3d_plain.so run plain
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r01w_probe_engine.jsonl') if l.startswith('{')]
for r in rows: print(r['tag'], r['op'], r['ms'], r['frac_hbm'])
PY
tail -c 300 gpurun_out/r01w.err
