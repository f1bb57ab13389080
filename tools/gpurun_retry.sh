#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <log> <command...>   (retries while the pod answers "transient"/busy)
to=$1; log=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|rc=3\|status=busy" $log; then sleep 150; continue; fi
  break
done
tail -60 $log
