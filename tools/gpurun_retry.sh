#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3: nothing charged).  usage: gpurun_retry.sh <timeout> <command>
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
