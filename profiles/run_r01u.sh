#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r01u_probe_engine.jsonl
run() { timeout 300 python tests/tools/probe_engine.py 24 $1 2>> gpurun_out/r01u.err | grep "p_sample shared" >> gpurun_out/r01u_probe_engine.jsonl; }
run default
for v in pss5o1 pss4o1 pss4; do SO3D_LIB_PATH=$PWD/build/variants/libso3d_$v.so run $v; done
run default
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r01u_probe_engine.jsonl') if l.startswith('{')]
for r in rows: print(r['tag'], r['op'], r['ms'], r['frac_hbm'])
PY
tail -c 300 gpurun_out/r01u.err
