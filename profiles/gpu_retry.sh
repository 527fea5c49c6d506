#!/bin/bash
# usage: profiles/gpu_retry.sh <timeout> '<command>'   -- retries gpurun while the pod answers busy (exit 3 / transient)
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 120; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
