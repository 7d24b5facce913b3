#!/bin/bash
# usage: scripts/gpurun_retry.sh <max_tries> <gpurun args...>   — retries while the pod answers "transient"/busy
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -90
  if echo "$out" | grep -q "status=transient\|status=busy\|status=refused"; then
    echo "[retry $i] not served, sleeping"; sleep 150; continue
  fi
  exit 0
done
exit 3
