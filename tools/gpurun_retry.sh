#!/bin/bash
# Developer: retry a gpurun call while the pod answers "transient" (exit code 3).  usage: tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 150
done
exit 3
