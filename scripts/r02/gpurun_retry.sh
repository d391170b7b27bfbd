#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers "transient" (nothing charged)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 150; continue; fi
  break
done
echo "attempts: $attempt" >> "$log"
