#!/bin/bash
# gpurun with retries while the pod answers "transient/busy" (exit code 3): scripts/gpurun_retry.sh [gpurun args...] -- 'cmd'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_retry.out 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" /tmp/gpurun_retry.out || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
cat /tmp/gpurun_retry.out
exit $rc
