#!/bin/bash
# usage: tools/gpu_retry.sh [gpurun options] -- 'command'
# Re-submits a gpurun call while the pod answers "busy" (exit code 3: nothing charged), up to ~40 minutes.
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 170
done
exit 3
