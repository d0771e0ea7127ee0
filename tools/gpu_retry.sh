#!/bin/bash
# developer tool: gpurun with retries while the pod answers "transient" (nothing is charged for those)
#   tools/gpu_retry.sh <gpus> <timeout_s> <script> <log>
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "bash $3" > "$4" 2>&1
  if grep -q "status=transient" "$4"; then sleep 150; else break; fi
done
