#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged).  Usage: gpurun_retry.sh <gpurun args...>
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt: no slot, sleeping 150 s" >&2
  sleep 150
done
exit 3
