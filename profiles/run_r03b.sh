#!/bin/bash
# r03b: (1) A/B of series-kernel builds (guard form x source order x block), (2) compute-sanitizer over every kernel family
mkdir -p gpurun_out
T=r03b
timeout 900 python tests/tools/probe_series.py 24 > gpurun_out/${T}_series_variants.jsonl 2> gpurun_out/${T}.err; cat gpurun_out/${T}_series_variants.jsonl | cut -c1-260
timeout 120 python tests/tools/sanitize_target.py 257 > gpurun_out/${T}_target_plain.log 2>&1; tail -2 gpurun_out/${T}_target_plain.log
for tool in memcheck racecheck synccheck initcheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python tests/tools/sanitize_target.py > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done|hazard|Error" gpurun_out/${T}_sanitizer_$tool.log | sort | uniq -c | head -12
done
