#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <command...>   — retries while the pod answers busy (rc 3 / transient)
T=$1; shift
for i in $(seq 1 12); do
  out=$(gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 120; continue; fi
  echo "$out"; exit $rc
done
echo "gave up"; exit 3
