#!/bin/bash
# r01p: L0 maps on the lean primitives -- GPU tests + engine probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01p_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01p_pytest.log
tail -15 gpurun_out/r01p_pytest.log
timeout 300 python tests/tools/probe_engine.py 24 default > gpurun_out/r01p_probe_engine.jsonl 2> gpurun_out/r01p.err
python - <<'PY'
import json
for l in open('gpurun_out/r01p_probe_engine.jsonl'):
    if l.startswith('{'):
        r=json.loads(l); print(r['op'].ljust(22), r.get('ms'), r.get('frac_hbm'))
PY
tail -c 300 gpurun_out/r01p.err
