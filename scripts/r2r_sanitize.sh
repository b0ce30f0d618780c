#!/bin/bash
# compute-sanitizer over the hot path (scripts/sanitize_cases.py): memcheck, racecheck, synccheck; logs -> gpurun_out/r2r_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/r2r_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r2r_$tool.log
  grep -E "^CASE|ERROR SUMMARY|RACECHECK SUMMARY|rc=" gpurun_out/r2r_$tool.log | tail -12
done
