#!/bin/bash
# r01k: warp-autonomous row engine -- GPU tests, then A/B of the engine schedules (probe + bench extras)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01k_pytest.log
tail -5 gpurun_out/r01k_pytest.log
timeout 300 python tests/tools/probe_engine.py 24 warp > gpurun_out/r01k_probe_engine.jsonl 2> gpurun_out/r01k_probe.err
SO3D_ENGINE=cta timeout 300 python tests/tools/probe_engine.py 24 cta >> gpurun_out/r01k_probe_engine.jsonl 2>> gpurun_out/r01k_probe.err
timeout 300 python tests/tools/probe_engine.py 24 warp >> gpurun_out/r01k_probe_engine.jsonl 2>> gpurun_out/r01k_probe.err
tail -c 400 gpurun_out/r01k_probe.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r01k_probe_engine.jsonl') if l.startswith('{')]
ops=[]
for r in rows:
    if r['op'] not in ops: ops.append(r['op'])
for o in ops:
    print(o.ljust(22), '  '.join(f"{r['tag']}:{r.get('ms','ERR')} ({r.get('frac_hbm','-')})" for r in rows if r['op']==o))
PY
