#!/bin/bash
# usage: scripts/gpu_retry.sh <gpus> <timeout> <script>   -- retries while the pod answers "busy / transient" (nothing is charged for those)
G=$1; T=$2; S=$3
for try in $(seq 1 20); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  echo "$out" | tail -14
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3\|retry in a few minutes"; then echo "[gpu_retry] try $try: busy, sleeping"; sleep 150; continue; fi
  break
done
