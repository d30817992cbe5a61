#!/bin/bash
# compute-sanitizer over one small pass of every kernel (scripts/run_small.py); summaries go to gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/run_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
