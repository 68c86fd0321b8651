#!/bin/bash
# r01m: forward-noising variants (register cap / output stages) + ncu capture of the warp-scheduled kernel
mkdir -p gpurun_out
: > gpurun_out/r01m_probe_engine.jsonl
for v in default qs4 qs4o2 qs5o2 qs6 default; do
  if [ $v = default ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$PWD/build/variants/libso3d_$v.so; fi
  timeout 300 python tests/tools/probe_engine.py 24 $v 2>> gpurun_out/r01m.err | grep "q_sample\|p_sample" >> gpurun_out/r01m_probe_engine.jsonl
done
unset SO3D_LIB_PATH
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r01m_probe_engine.jsonl') if l.startswith('{')]
ops=[]
for r in rows:
    if r['op'] not in ops: ops.append(r['op'])
for o in ops:
    print(o.ljust(22), '  '.join(f"{r['tag']}:{r.get('ms','ERR')} ({r.get('frac_hbm','-')})" for r in rows if r['op']==o))
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:QSampleOp' -s 4 -c 1 -f -o gpurun_out/r01m_prof_qsample \
    python tests/tools/probe_one.py q_sample 22 > gpurun_out/r01m_ncu_stdout.log 2>&1
tail -c 300 gpurun_out/r01m.err; tail -3 gpurun_out/r01m_ncu_stdout.log
