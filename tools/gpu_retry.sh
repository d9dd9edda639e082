#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <command string>  - retries gpurun while the pod answers busy (exit 3)
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
