#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout_s> '<command>'  — retries while the pod answers "transient/busy"
cd "$(dirname "$0")/.."
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@" 2>&1)
  if ! echo "$out" | grep -q "status=transient"; then echo "$out"; exit 0; fi
  sleep 120
done
echo "$out"
