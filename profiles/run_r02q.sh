#!/bin/bash
# r02q: final check at HEAD (series kernel: source order 1, 64-term blocks, strength-reduced tile indices) -- GPU tests, smoke, bench line, reference arm
mkdir -p gpurun_out
[ -s gpurun_out/r02q_pytest.log ] || { timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_pytest.log; }
tail -4 gpurun_out/r02q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02q_smoke.log 2>&1; tail -1 gpurun_out/r02q_smoke.log
SECONDS=0; timeout 900 python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench wall ${SECONDS} s"
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02q_bench_reference.json 2>> gpurun_out/r02q_bench.err
wc -l gpurun_out/r02q_bench.json gpurun_out/r02q_bench_reference.json
