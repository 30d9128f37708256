#!/usr/bin/env bash
# gpurun with retries on "busy" (exit 3: nothing charged).  Usage: tools/gpurun_retry.sh <timeout-seconds> <log> <command...>
T=$1; LOG=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
