#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: scripts/gpurun_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
