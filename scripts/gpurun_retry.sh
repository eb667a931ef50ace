#!/bin/bash
# Local helper: keep asking for a GPU box until the call is accepted (exit codes 2 = refused, 3 = busy are retried
# for busy only).  Usage: scripts/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then sleep 150; continue; fi
  exit $rc
done
exit 3
