#!/bin/bash
# usage (CPU box): bash scripts/gpurun_retry.sh <log> <gpurun args...>  -- retries while the pod answers "transient" / busy
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$LOG" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  exit $rc
done
exit 3
