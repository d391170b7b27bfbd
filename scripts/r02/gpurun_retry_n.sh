#!/bin/bash
# usage: gpurun_retry_n.sh <gpus> <log> <timeout> <command...>   -- multi-GPU variant of gpurun_retry.sh
gpus=$1; shift; log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$gpus" --timeout "$to" -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|rc=3\|no box\|busy" "$log" && ! grep -q "status=ok" "$log"; then sleep 150; continue; fi
  break
done
echo "attempts: $attempt" >> "$log"
