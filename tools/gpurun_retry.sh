#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   retries while the pod answers "transient" (nothing charged)
log=$1; shift
for try in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 150; else break; fi
done
tail -40 "$log"
