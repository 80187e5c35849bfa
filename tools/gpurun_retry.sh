#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>   — retries while the pod answers "busy"
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -80 /tmp/gpurun_last.log
exit $rc
