#!/bin/bash
# developer helper: submit a session script to the GPU box, retrying while the pod answers "busy" (nothing is charged then)
# usage: tools/gpu_sessions/submit.sh <timeout-seconds> <script> [gpus]
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(gpurun --timeout $T -- "bash $S" 2>&1); else out=$(gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  if echo "$out" | grep -q "status=transient\|status=busy\|retry in a few minutes"; then sleep 90; continue; fi
  echo "$out" | tail -25
  exit 0
done
echo "gave up after 40 tries"; echo "$out" | tail -5
