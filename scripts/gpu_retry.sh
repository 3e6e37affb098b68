#!/bin/bash
# usage: gpu_retry.sh <timeout_s> [--gpus N] <command>   — retries while the pod answers busy (exit 3 / transient)
t=$1; shift
opts=""
if [ "$1" == "--gpus" ]; then opts="--gpus $2"; shift 2; fi
for i in 1 2 3 4 5 6 7 8 9 10; do
  out=$(gpurun --timeout $t $opts -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|retry later\|busy"; then sleep 150; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
