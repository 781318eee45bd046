#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script>   -- retries while the pod answers "busy" (nothing charged)
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out" | tail -40
  exit 0
done
echo "gave up: pod busy"
