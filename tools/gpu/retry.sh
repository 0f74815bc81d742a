#!/bin/bash
# retry a gpurun call while the pod answers "transient" (busy): tools/gpu/retry.sh <log file> <timeout> <command...>
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then break; fi
  sleep 45
done
