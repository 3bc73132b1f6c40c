#!/bin/bash
# compute-sanitizer over a reduced-size pass of every kernel (scripts/sanitize_target.py).  ONE tool per gpurun call
# (B200_PROFILING.md: several tools in one call have left a GPU unusable):
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh memcheck'      # then racecheck, synccheck, initcheck
# Writes gpurun_out/sanitize_<tool>.log; the summary lines are copied to profiles/r2_sanitize.txt by hand.
set -u
tool=${1:-memcheck}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 0 python scripts/sanitize_target.py > gpurun_out/sanitize_${tool}.log 2>&1
echo "exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|========= (Invalid|Race|Barrier|Uninit|Error)" gpurun_out/sanitize_${tool}.log | sort | uniq -c | head -40
