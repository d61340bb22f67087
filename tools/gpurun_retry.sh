#!/bin/bash
# tools/gpurun_retry.sh <timeout-seconds> <command...>: gpurun with retries while the pod answers "busy / draining" (exit code 3);
# GPUS=N in the environment asks for N GPUs of one box
T=$1; shift
G=${GPUS:+--gpus $GPUS}
for n in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T $G -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc = 3 ]; then sleep 90; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
exit $rc
