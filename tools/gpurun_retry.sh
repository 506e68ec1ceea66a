#!/bin/bash
# usage: [GPURUN_FLAGS="--gpus 2"] tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun exit $rc" >> $log; exit $rc; fi
  sleep 90
done
