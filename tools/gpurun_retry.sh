#!/bin/bash
# gpurun with retries while the pod answers "busy / transient" (exit 3, nothing charged).  Usage: tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
