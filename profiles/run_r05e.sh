#!/bin/bash
# r05e: final GPU tests + sanitizers at HEAD (after the SE(3) per-row-t step moved to the two-row adaptor)
mkdir -p gpurun_out
T=r05e
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
for tool in racecheck synccheck memcheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 10 python tests/tools/sanitize_target.py 1 257 1300 > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done" gpurun_out/${T}_sanitizer_$tool.log | sort | uniq -c | head -4
done
