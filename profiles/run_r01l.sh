#!/bin/bash
# r01l: per-op engine schedule + paired-term series -- GPU tests, A/B of build variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r01l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01l_pytest.log
tail -5 gpurun_out/r01l_pytest.log
: > gpurun_out/r01l_series_ab.jsonl
for v in default pairs0 default pairs0; do
  if [ $v = default ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$PWD/build/variants/libso3d_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --no-e2e 2>> gpurun_out/r01l.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({'variant':'$v','evals_per_s':d['value'],'ms':d['ms_per_step'],'clk_per_warp_term':d['roofline']['clk_per_warp_term'],'frac':d['roofline']['frac']}))
" | tee -a gpurun_out/r01l_series_ab.jsonl
done
unset SO3D_LIB_PATH
timeout 300 python tests/tools/probe_engine.py 24 default > gpurun_out/r01l_probe_engine.jsonl 2>> gpurun_out/r01l.err
SO3D_LIB_PATH=$PWD/build/variants/libso3d_nocap.so timeout 300 python tests/tools/probe_engine.py 24 nocap >> gpurun_out/r01l_probe_engine.jsonl 2>> gpurun_out/r01l.err
timeout 300 python tests/tools/probe_engine.py 24 default >> gpurun_out/r01l_probe_engine.jsonl 2>> gpurun_out/r01l.err
tail -c 400 gpurun_out/r01l.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r01l_probe_engine.jsonl') if l.startswith('{')]
ops=[]
for r in rows:
    if r['op'] not in ops: ops.append(r['op'])
for o in ops:
    print(o.ljust(22), '  '.join(f"{r['tag']}:{r.get('ms','ERR')} ({r.get('frac_hbm','-')})" for r in rows if r['op']==o))
PY
