#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (exit code 3): usage gpurun_retry.sh <log> [gpurun args...]
log=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log" && [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
