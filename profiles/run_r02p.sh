#!/bin/bash
# r02p: record run at HEAD (round-1 final; series kernel on source order 1 / block 64): GPU tests, smoke, bench line, reference arm, probe, launch list, series capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02p_pytest.log
tail -4 gpurun_out/r02p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02p_smoke.log 2>&1; tail -2 gpurun_out/r02p_smoke.log
timeout 900 python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; tail -c 600 gpurun_out/r02p_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02p_bench_reference.json 2>> gpurun_out/r02p_bench.err
timeout 300 python tests/tools/probe_engine.py 24 default > gpurun_out/r02p_probe_engine.jsonl 2>> gpurun_out/r02p_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02p_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02p_ncu_launches_stdout.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:LogpScoreOp -s 3 -c 1 -f -o gpurun_out/r02p_prof_series \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-extra --n 4194304 > gpurun_out/r02p_ncu_series_stdout.log 2>&1
ls -la gpurun_out | tail -12
