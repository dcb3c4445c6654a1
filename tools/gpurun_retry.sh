#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'   — retries while the pod answers "transient" (no box free)
T=$1; G=$2; shift 2
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] transient, sleeping"; sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3
