#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers busy/transient (nothing is charged for those)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|rc=3\|refused\|busy" "$log" && ! grep -q "status=ok" "$log"; then sleep 90; continue; fi
  break
done
