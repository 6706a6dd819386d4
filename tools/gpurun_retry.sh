#!/bin/bash
# usage: gpurun_retry.sh <timeout> <extra gpurun args...> -- <cmd>   : retries while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  break
done
