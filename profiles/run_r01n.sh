#!/bin/bash
# r01n: one-round-trip multi-point bucket lookup, forward noising at 4 CTAs -- GPU tests, probes, ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01n_pytest.log
tail -4 gpurun_out/r01n_pytest.log
: > gpurun_out/r01n_probe_engine.jsonl
for v in default ps4 qs5 default; do
  if [ $v = default ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$PWD/build/variants/libso3d_$v.so; fi
  timeout 300 python tests/tools/probe_engine.py 24 $v 2>> gpurun_out/r01n.err | grep "q_sample\|p_sample\|sample" >> gpurun_out/r01n_probe_engine.jsonl
done
unset SO3D_LIB_PATH
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r01n_probe_engine.jsonl') if l.startswith('{')]
ops=[]
for r in rows:
    if r['op'] not in ops: ops.append(r['op'])
for o in ops:
    print(o.ljust(22), '  '.join(f"{r['tag']}:{r.get('ms','ERR')} ({r.get('frac_hbm','-')})" for r in rows if r['op']==o))
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:QSampleOp' -s 4 -c 1 -f -o gpurun_out/r01n_prof_qsample \
    python tests/tools/probe_one.py q_sample 22 > gpurun_out/r01n_ncu_stdout.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:PStepOp' -s 4 -c 1 -f -o gpurun_out/r01n_prof_pstep \
    python tests/tools/probe_one.py p_sample 22 >> gpurun_out/r01n_ncu_stdout.log 2>&1
tail -c 300 gpurun_out/r01n.err; tail -2 gpurun_out/r01n_ncu_stdout.log
