#!/bin/bash
# retry.sh <timeout> [--gpus N] <command>: gpurun with retries while the pod answers busy (exit 3 / transient)
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 120; continue; fi
  echo "$out"; exit $rc
done
echo "gave up after 20 tries"; exit 3
