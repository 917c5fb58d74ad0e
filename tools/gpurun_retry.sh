#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> [--gpus N] <command...>   -- retries while the pod answers busy (exit 3)
log=$1; shift; to=$1; shift
extra=""
if [ "$1" == "--gpus" ]; then extra="--gpus $2"; shift; shift; fi
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
