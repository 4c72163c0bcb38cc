#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script>   — retries while the pod answers "transient" (nothing charged)
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" 2>&1)
  echo "$out" | tail -40
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  sleep 90
done
