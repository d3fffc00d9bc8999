#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python tools/sanitize_r2.py > $O/san_plain.log 2>&1; echo "plain rc $?"; tail -3 $O/san_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_r2.py > $O/san_$tool.log 2>&1; echo "$tool rc $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" $O/san_$tool.log | head -8
done
