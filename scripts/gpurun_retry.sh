#!/bin/bash
# usage: [GPUS=N] scripts/gpurun_retry.sh <logfile> <timeout> <command...>   (retries while the pod answers busy/transient)
log=$1; shift; to=$1; shift
extra=""; [ -n "${GPUS:-}" ] && extra="--gpus $GPUS"
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $extra --timeout "$to" -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|exit code 3" "$log"; then sleep 45; continue; fi
  break
done
