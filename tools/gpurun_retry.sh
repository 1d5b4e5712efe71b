#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <log> <command...>   -- resubmits while the pool answers "busy" (nothing charged)
t=$1; log=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "attempts=$attempt rc=$rc" >> $log
