#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT [--gpus N] -- 'command'   (retries while the pod answers busy / transient)
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry done rc=$rc" >> "$LOG"
