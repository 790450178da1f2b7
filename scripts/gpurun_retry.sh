#!/bin/bash
# usage: scripts/gpurun_retry.sh LOGFILE [gpurun args...] -- keeps retrying while the pod answers busy / transient
# (exit 3, or a "status=transient" verdict: nothing was charged)
LOG=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 60
done
exit 3
