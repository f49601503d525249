#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout_s> <command...>  -- retries while the pod answers "busy" (exit 3, nothing charged)
log=$1; to=$2; shift 2
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then echo "exit $rc" >> "$log"; exit $rc; fi
  sleep 45
done
echo "exit 3 (gave up)" >> "$log"; exit 3
