#!/bin/bash
# one compute-sanitizer tool per gpurun call: scripts/r02_sanitize.sh memcheck|racecheck
tool=${1:-memcheck}
out=gpurun_out
python scripts/sanitize_case.py > $out/r02_sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/r02_sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py > $out/r02_sanitize_$tool.log 2>&1
echo "== $tool rc $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|bit-identical|streaming" $out/r02_sanitize_$tool.log | tail -20
