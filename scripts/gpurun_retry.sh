#!/bin/bash
# Runs one gpurun call, retrying while the pod answers "busy" (exit code 3; nothing is charged for those).
# usage: scripts/gpurun_retry.sh LOGFILE [gpurun args...]
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] attempt $attempt finished rc=$rc" >> "$log"; exit $rc; fi
  sleep 120
done
echo "[retry] gave up" >> "$log"; exit 3
