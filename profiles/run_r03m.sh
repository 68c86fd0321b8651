#!/bin/bash
# r03m: warp-split small-batch series + two-row reverse step at HEAD: GPU tests, smoke, full bench, reference arm
mkdir -p gpurun_out
T=r03m
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -8 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -c 300 gpurun_out/${T}_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r03m_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "acc", d["accuracy"]["max_rel_err_f"], d["accuracy"]["max_rel_err_score"])
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "clocks", d["clocks"])
    for k in ("series_small_batch", "reverse_particle_steps_per_sec", "noised_rotations_per_sec", "noised_rotations_with_score_per_sec", "reverse_loop_1000_steps", "size_sweep"):
        print(k, json.dumps(d["extra"][k])[:1200])
except Exception as e:
    print("no bench line:", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
