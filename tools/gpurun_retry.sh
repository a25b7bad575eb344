#!/bin/bash
# usage: tools/gpurun_retry.sh <tries> <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
TRIES=$1; shift
for i in $(seq 1 $TRIES); do
  OUT=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] pod busy"; sleep 90; continue; fi
  echo "$OUT" | tail -40
  exit 0
done
echo "gave up after $TRIES tries"
