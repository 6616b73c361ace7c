#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers "busy" (exit code 3)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> "$LOG"; echo finished >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$LOG"; echo finished >> "$LOG"
