#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT [--gpus N] -- command...   (retries while the pod has no free slot)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 90; continue; fi
  break
done
