#!/bin/sh
# usage: tools/gpurun_retry.sh <logfile> <timeout_s> [--gpus N] '<command>'   -- retries while the pod answers busy (exit 3 / transient)
LOG=$1; TMO=$2; shift 2
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout "$TMO" -- "$@" > "$LOG" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 60
done
exit 3
