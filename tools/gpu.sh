#!/bin/bash
# Developer helper: gpurun with retries while the pod answers "busy" (exit 3).   tools/gpu.sh TIMEOUT 'command'
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $T -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
