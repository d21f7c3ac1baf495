#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy (nothing is charged for those)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|rc=3\|busy" $log && ! grep -q "charged=[1-9]" $log; then sleep 120; continue; fi
  break
done
