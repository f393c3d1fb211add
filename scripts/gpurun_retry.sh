#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <script> [gpus]   -- retries while the pod answers "busy" (nothing charged)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then OUT=$(gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(gpurun --gpus $G --timeout $T -- "bash $S $G" 2>&1); fi
  echo "$OUT" | tail -80
  if echo "$OUT" | grep -q "status=transient\|status=busy\|nothing was charged"; then sleep 150; continue; fi
  break
done
