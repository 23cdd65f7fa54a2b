#!/bin/bash
# Run every GPU test function in its own process under a hard timeout, so that one hung kernel cannot eat the
# whole GPU lease.  Usage: scripts/run_gpu_tests_isolated.sh [per-function timeout seconds] [log]
T=${1:-150}
LOG=${2:-gpurun_out/pytest_gpu_isolated.log}
: > "$LOG"
funcs=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
pass=0; fail=0
for f in $funcs; do
  start=$(date +%s)
  timeout -k 5 "$T" python -m pytest "$f" -m gpu -q -x --no-header -p no:cacheprovider >> "$LOG" 2>&1
  rc=$?
  dt=$(( $(date +%s) - start ))
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS ${dt}s $f"; else fail=$((fail+1)); echo "FAIL(rc=$rc) ${dt}s $f"; fi
done
echo "isolated gpu tests: $pass function(s) passed, $fail failed"
