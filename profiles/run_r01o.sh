#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r01o_probe_engine.jsonl
run() { timeout 300 python tests/tools/probe_engine.py 24 $1 2>> gpurun_out/r01o.err | grep "q_sample\|p_sample\|sample\|auto" >> gpurun_out/r01o_probe_engine.jsonl; }
run default
SO3D_ENGINE=warp run warp
SO3D_LIB_PATH=$PWD/build/variants/libso3d_pss4.so SO3D_ENGINE=warp run pss4-warp
SO3D_LIB_PATH=$PWD/build/variants/libso3d_pss4.so SO3D_ENGINE=cta run pss4-cta
run default
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r01o_probe_engine.jsonl') if l.startswith('{')]
ops=[]
for r in rows:
    if r['op'] not in ops: ops.append(r['op'])
for o in ops:
    print(o.ljust(22), '  '.join(f"{r['tag']}:{r.get('ms','ERR')} ({r.get('frac_hbm','-')})" for r in rows if r['op']==o))
PY
tail -c 300 gpurun_out/r01o.err
