#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit 3, nothing charged).
#   tools/gpurun_retry.sh [gpurun args...] -- 'command'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] attempt $attempt busy; sleeping 120 s"
  sleep 120
done
exit 3
