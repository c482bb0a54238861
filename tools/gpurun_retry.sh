#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command...> — retries while the pod answers "transient"/busy
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); echo "$out" | tail -n 25
  if echo "$out" | grep -q "status=transient\|answers busy\|retry in a few minutes"; then sleep 100; continue; fi
  break
done
