#!/bin/bash
# gpurun with retries on "no box / slot free right now" (exit code 3): tools/gpu_retry.sh [gpurun args...] -- '<command>'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
