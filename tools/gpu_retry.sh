#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <gpurun args...>   - retries while the pod answers "busy" (exit code 3 / transient)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 150; continue; fi
  break
done
