#!/bin/bash
# r02c: first tile loads issued before the op setup -- tests, probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -4 gpurun_out/r02c_pytest.log
timeout 300 python tests/tools/probe_engine.py 24 default > gpurun_out/r02c_probe_engine.jsonl 2> gpurun_out/r02c.err
timeout 300 python tests/tools/probe_engine.py 24 default >> gpurun_out/r02c_probe_engine.jsonl 2>> gpurun_out/r02c.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r02c_probe_engine.jsonl') if l.startswith('{')]
ops=[]
for r in rows:
    if r['op'] not in ops: ops.append(r['op'])
for o in ops:
    print(o.ljust(22), '  '.join(f"{r.get('ms','ERR')} ({r.get('frac_hbm','-')})" for r in rows if r['op']==o))
PY
tail -c 300 gpurun_out/r02c.err
