#!/bin/bash
# gpurun with retries while the pod answers "transient" (busy): tools/gpurun_retry.sh <log> <timeout> <command...>
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then break; fi
  sleep 90
done
tail -5 $log
