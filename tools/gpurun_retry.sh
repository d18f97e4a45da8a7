#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pod answers "transient"/busy)
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|status=busy"; then
    echo "[retry] attempt $attempt transient; sleeping 150s"
    sleep 150
    continue
  fi
  break
done
