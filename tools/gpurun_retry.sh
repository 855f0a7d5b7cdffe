#!/usr/bin/env bash
# Developer helper: gpurun with retries while the pod has no free slot (exit 3 / "draining", nothing charged).
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout_s> '<command>'
t="$1"; shift
for i in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$t" -- "$@" 2>&1); rc=$?
    if echo "$out" | grep -qE "nothing was charged|retry in [0-9]+s"; then sleep 60; continue; fi
    echo "$out" | tail -4
    exit $rc
done
echo "gpurun_retry: gave up"; exit 3
