#!/bin/bash
# r05n: one-launch reverse process, lookups straight-line + one warp vote (lib) vs per-row branches (novote)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "one_launch or loop or reverse_process" 2>&1 | tail -2
for v in novote "" novote ""; do
  if [ -z "$v" ]; then lib=""; tag=vote; else lib=build/variants/libso3d_$v.so; tag=$v; fi
  SO3D_LIB_PATH=$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-sweep --no-eager --no-accuracy 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); v = d['extra']['reverse_loop_1000_steps']
print(json.dumps({'tag': '$tag', 'one_launch_s': v['one_launch']['seconds'], 'per_step_s': v['seconds']}))" >> gpurun_out/r05n_probe.txt
done
cat gpurun_out/r05n_probe.txt
