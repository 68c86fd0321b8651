#!/bin/bash
# r03c: loop kernel + draw diet + sanitizer follow-up.  GPU tests, series variants (round 2), fused-kernel A/B, racecheck on the
# row engines after the init barrier, synccheck per denoiser section, full bench.
mkdir -p gpurun_out
T=r03c
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
timeout 900 python tests/tools/probe_series.py 24 > gpurun_out/${T}_series_variants.jsonl 2> gpurun_out/${T}.err
python - <<'PY'
import json
for l in open("gpurun_out/r03c_series_variants.jsonl"):
    d = json.loads(l); print(d.get("lib"), d.get("series_ms"), d.get("series_pure_ms"), d.get("series_adaptive_ms"), d.get("series_sum", [0])[0], d.get("error", ""))
PY
for lib in shipped build/variants_fused/libso3d_nomufu.so build/variants_fused/libso3d_qsx3.so; do
  if [ "$lib" = shipped ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$lib; fi
  timeout 300 python tests/tools/probe_engine.py 24 $(basename $lib .so) >> gpurun_out/${T}_probe_engine.jsonl 2>> gpurun_out/${T}.err
done
unset SO3D_LIB_PATH
python - <<'PY'
import json, collections
t = collections.defaultdict(dict)
for l in open("gpurun_out/r03c_probe_engine.jsonl"):
    d = json.loads(l); t[d["op"]][d["tag"]] = (d.get("ms"), d.get("frac_hbm"))
for op, v in t.items(): print(op, v)
PY
timeout 400 compute-sanitizer --tool racecheck --print-limit 10 python tests/tools/sanitize_target.py rows 257 1300 > gpurun_out/${T}_sanitizer_racecheck_rows.log 2>&1
echo "== racecheck rows exit $?"; grep -E "RACECHECK SUMMARY|sanitize target done" gpurun_out/${T}_sanitizer_racecheck_rows.log
for sec in rows denoiser_step denoiser_loop1 denoiser_loop5; do
  timeout 400 compute-sanitizer --tool synccheck --print-limit 5 python tests/tools/sanitize_target.py $sec 257 > gpurun_out/${T}_sanitizer_synccheck_$sec.log 2>&1
  echo "== synccheck $sec exit $?"; grep -E "ERROR SUMMARY|sanitize target done|denoiser n|Barrier error" gpurun_out/${T}_sanitizer_synccheck_$sec.log | sort | uniq -c | head -8
done
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/${T}_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r03c_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "acc", d["accuracy"]["max_rel_err_f"], d["accuracy"]["max_rel_err_score"])
    for k in ("reverse_particle_steps_per_sec", "noised_rotations_per_sec", "noised_rotations_with_score_per_sec", "reverse_loop_1000_steps", "se3_frames_cfg5"):
        print(k, json.dumps(d["extra"][k])[:900])
except Exception as e:
    print("no bench line:", e)
PY
