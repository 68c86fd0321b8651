#!/bin/bash
# r04x: record lookup "fast path first, search overrides" (new) vs branch first (base = HEAD): tests, A/B incl. the one-launch loop
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in base "" base ""; do
  if [ -z "$v" ]; then lib=""; tag=new; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python tests/tools/probe_engine.py 24 $tag 2>&1 | grep -E "sample" >> gpurun_out/r04x_probe.txt
  SO3D_LIB_PATH=$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-sweep --no-eager --no-accuracy 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); v = d['extra']['reverse_loop_1000_steps']
print(json.dumps({'tag': '$tag', 'op': 'one-launch loop s', 'ms': v['one_launch']['seconds'], 'frac_hbm': v['seconds']}))" >> gpurun_out/r04x_probe.txt
done
cut -c1-175 gpurun_out/r04x_probe.txt
