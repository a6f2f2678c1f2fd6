#!/bin/bash
# gpurun with retries while the pod answers "transient" (draining / no slot): bash profiles/scripts/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_try.log 2>&1
  if grep -q "status=transient" /tmp/gpurun_try.log; then sleep 150; continue; fi
  break
done
tail -60 /tmp/gpurun_try.log
