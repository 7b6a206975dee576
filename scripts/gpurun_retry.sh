#!/bin/bash
# usage: gpurun_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers busy (exit 3), at most 40 times
out=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] finished rc=$rc after $i attempt(s)" >> "$out"; exit $rc; fi
  sleep 90
done
echo "[retry] gave up" >> "$out"; exit 3
