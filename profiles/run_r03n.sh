#!/bin/bash
# r03n: record run at HEAD (round 2): GPU tests, smoke, bench line, reference arm, launch list, ncu captures of the kernels that
# changed after r03d (two-row reverse step, warp-split series)
mkdir -p gpurun_out
T=r03n
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r03n_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "acc", d["accuracy"]["max_rel_err_f"], d["accuracy"]["max_rel_err_score"])
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "clocks", d["clocks"])
    print("small", json.dumps(d["extra"]["series_small_batch"])[:600])
    r = json.load(open("gpurun_out/r03n_bench_reference.json")); print("reference", r["value"], "e2e ratio", d["e2e"]["value"] / r["value"], "ratio", d["value"] / r["value"])
except Exception as e:
    print("no bench line:", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sweep --no-eager --no-accuracy > gpurun_out/${T}_ncu_launches_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:PStep2Op -s 4 -c 1 -f -o gpurun_out/${T}_prof_pstep2 \
    python tests/tools/probe_one.py p_sample 22 > gpurun_out/${T}_ncu_pstep2_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:series_warp_kernel -s 2 -c 1 -f -o gpurun_out/${T}_prof_series_warp \
    python tests/tools/probe_one.py series_small 12 > gpurun_out/${T}_ncu_series_warp_stdout.log 2>&1
ls -la gpurun_out | grep ${T}
