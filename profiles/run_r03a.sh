#!/bin/bash
# r03a (round 2, first GPU call): series guard + bench rework.  GPU tests, smoke, full bench line (accuracy, size sweep,
# cfg 4 / cfg 5 legs, torch-CUDA-eager arm, CPU legs), reference arm.
mkdir -p gpurun_out
T=r03a
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"; tail -c 1500 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r03a_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"])
    print("accuracy", d["accuracy"])
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print("clocks", d["clocks"])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:700])
    print("cpu", json.dumps(d.get("cpu_baseline"))[:900])
except Exception as e:
    print("no bench line:", e)
PY
