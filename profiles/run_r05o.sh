#!/bin/bash
# r05o: final GPU tests + smoke + sanitizers + bench line + reference arm at HEAD
mkdir -p gpurun_out
T=r05o
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
for tool in racecheck synccheck memcheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 10 python tests/tools/sanitize_target.py 1 257 1300 > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done" gpurun_out/${T}_sanitizer_$tool.log | sort | uniq -c | head -4
done
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r05o_bench.json 2> gpurun_out/r05o_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r05o_bench_reference.json 2>> gpurun_out/r05o_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r05o_bench.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "acc", d["accuracy"]["max_rel_err_f"], d["accuracy"]["max_rel_err_score"])
for k in ("score_evals_per_sec_auto", "reverse_particle_steps_per_sec", "noised_rotations_per_sec", "noised_rotations_with_score_per_sec"):
    v = d["extra"][k]; print(k, v.get("value"), v["roofline"]["frac"])
v = d["extra"]["se3_frames_cfg5"]; print("se3", v["noising_frames_per_sec"]["value"], v["noising_frames_per_sec"]["roofline"]["frac"], v["reverse_frame_steps_per_sec"]["value"], v["reverse_frame_steps_per_sec"]["roofline"]["frac"])
v = d["extra"]["reverse_loop_1000_steps"]; print("loop", v["seconds"], v["one_launch"]["seconds"])
r = json.load(open("gpurun_out/r05o_bench_reference.json")); print("reference", r["value"], "e2e ratio", d["e2e"]["value"] / r["value"])
PY
