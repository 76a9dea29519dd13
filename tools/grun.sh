#!/bin/bash
# gpurun with retries on "no box / slot free" answers (nothing is charged for those).  Usage: tools/grun.sh TIMEOUT 'command'
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|retry in a few minutes"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
