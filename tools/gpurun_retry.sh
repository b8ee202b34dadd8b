#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'  [extra gpurun args]
# retries while gpurun answers "no box / slot free" (exit code 3); every other outcome is final
t=$1; shift
cmd=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" -- "$cmd"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
