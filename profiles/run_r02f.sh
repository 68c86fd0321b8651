#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tools/probe_engine.py 24 default 2> gpurun_out/r02f.err | grep "se3\|p_sample\|q_sample" | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['op'].ljust(26), r.get('ms'), r.get('frac_hbm'), r.get('error'))"
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/r02f_bench.json 2>> gpurun_out/r02f.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print(d['value'])
for k in ('reverse_loop_20000_particles_ms','train_step_batch256_ms','mmd_pairs_per_sec'): print(k, d['extra'].get(k))
PY
tail -c 600 gpurun_out/r02f.err
