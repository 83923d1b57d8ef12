#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout_s> '<command>'   — retries while gpurun answers "busy" (exit 3)
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "rc=$rc" >> "$log"; exit $rc; fi
  sleep 45
done
echo "gave up" >> "$log"
