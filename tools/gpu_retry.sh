#!/bin/bash
# usage: tools/gpu_retry.sh LOGNAME TIMEOUT [--gpus N] -- 'command'   (retries while the pod answers busy)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > "gpurun_out/$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc attempt=$attempt" >> "gpurun_out/$log"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "gpurun_out/$log"
