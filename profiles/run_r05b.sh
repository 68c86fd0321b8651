#!/bin/bash
# r05b: where do the 5 % of the per-step-launch reverse loop go in a full bench run? (legs switched off one at a time)
mkdir -p gpurun_out
for flags in "" "--no-cpu" "--no-e2e" "--no-cpu --no-e2e"; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-sweep --no-eager --no-accuracy $flags 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); v = d['extra']['reverse_loop_1000_steps']
print(json.dumps({'flags': '$flags', 'per_step_loop_s': v['seconds'], 'one_launch_s': v['one_launch']['seconds'], 'p_step_ms': d['extra']['reverse_particle_steps_per_sec']['ms_per_step']}))" >> gpurun_out/r05b_probe.txt
done
cat gpurun_out/r05b_probe.txt
