#!/bin/bash
# gpurun with retries while the pod answers "busy" (nothing is charged for those): scripts/gpurun_retry.sh [gpurun args] -- cmd
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out" | tail -60
  if ! echo "$out" | grep -q "status=transient"; then exit $rc; fi
  echo "[retry] attempt $attempt was transient; sleeping 90 s"; sleep 90
done
exit 3
