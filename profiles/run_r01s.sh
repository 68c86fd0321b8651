#!/bin/bash
# r01s: engine index-math diet -- tests, probe, dynamic instruction counts
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01s_pytest.log
tail -4 gpurun_out/r01s_pytest.log
timeout 300 python tests/tools/probe_engine.py 24 default > gpurun_out/r01s_probe_engine.jsonl 2> gpurun_out/r01s.err
timeout 300 python tests/tools/probe_engine.py 24 default >> gpurun_out/r01s_probe_engine.jsonl 2>> gpurun_out/r01s.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r01s_probe_engine.jsonl') if l.startswith('{')]
ops=[]
for r in rows:
    if r['op'] not in ops: ops.append(r['op'])
for o in ops:
    print(o.ljust(22), '  '.join(f"{r.get('ms','ERR')} ({r.get('frac_hbm','-')})" for r in rows if r['op']==o))
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:PStepOp' -s 4 -c 1 -f -o gpurun_out/r01s_prof_pstep \
    python tests/tools/probe_one.py p_sample 22 > gpurun_out/r01s_ncu_stdout.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:QSampleOp' -s 4 -c 1 -f -o gpurun_out/r01s_prof_qsample \
    python tests/tools/probe_one.py q_sample 22 >> gpurun_out/r01s_ncu_stdout.log 2>&1
tail -c 300 gpurun_out/r01s.err
