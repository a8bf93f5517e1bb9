#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (scripts/sanitize_cases.py): memcheck, racecheck
# (shared-memory hazards: the kernels order their shared-memory exchanges with __syncwarp / named barriers / mbarriers) and
# synccheck. Summaries land in gpurun_out/ (copied to profiles/ by hand).
# Usage: gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
set -u
O=gpurun_out
mkdir -p $O
python scripts/sanitize_cases.py > $O/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py > $O/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|^\{" $O/sanitize_$tool.log | tail -6
done
