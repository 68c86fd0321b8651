#!/bin/bash
# r04t: record run at HEAD (round 2: + shared-memory float-format guide, two-row closed-form score, TwoRow adaptor): GPU tests, smoke, bench line (driver's flags), reference arm,
# launch list, sanitizers over every kernel family once more
mkdir -p gpurun_out
T=r04t
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r04t_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "acc", d["accuracy"]["max_rel_err_f"], d["accuracy"]["max_rel_err_score"])
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "clocks", d["clocks"])
    for k in ("reverse_particle_steps_per_sec", "noised_rotations_per_sec", "noised_rotations_with_score_per_sec", "se3_frames_cfg5", "reverse_loop_1000_steps"):
        v = d["extra"][k]; print(k, v.get("value"), json.dumps(v.get("roofline", v.get("noising_frames_per_sec")))[:200], json.dumps(v.get("one_launch", v.get("reverse_frame_steps_per_sec", "")))[:160])
    r = json.load(open("gpurun_out/r04t_bench_reference.json")); print("reference", r["value"], "e2e ratio", d["e2e"]["value"] / r["value"], "ratio", d["value"] / r["value"])
except Exception as e:
    print("no bench line:", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sweep --no-eager --no-accuracy > gpurun_out/${T}_ncu_launches_stdout.log 2>&1
for tool in racecheck synccheck memcheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 10 python tests/tools/sanitize_target.py 1 257 1300 > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done" gpurun_out/${T}_sanitizer_$tool.log | sort | uniq -c | head -4
done
