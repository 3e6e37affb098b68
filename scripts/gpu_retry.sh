#!/bin/bash
# usage: gpu_retry.sh <timeout_s> <command...>   — retries while the pod answers busy (exit 3 / transient)
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(gpurun --timeout $t -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|retry in a few minutes"; then sleep 120; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
