#!/bin/bash
# r03y: eps prefetched one tile ahead in the closed-form / auto score kernels: A/B vs HEAD (base), all ops of the probe
mkdir -p gpurun_out
for v in base "" base ""; do
  if [ -z "$v" ]; then lib=""; tag=new; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 >> gpurun_out/r03y_probe.txt
done
python - <<'PY'
import json
for l in open('gpurun_out/r03y_probe.txt'):
    try: d=json.loads(l)
    except Exception: continue
    print(d.get('tag'), d.get('op'), d.get('ms'), d.get('frac_hbm'))
PY
timeout 600 python -m pytest tests -m gpu -q -x -k "logp or score or igso3 or series" 2>&1 | tail -3
