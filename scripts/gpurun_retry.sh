#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): scripts/gpurun_retry.sh <gpurun args...>
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_last.log 2>&1
  if ! grep -q "status=transient" /tmp/gpurun_last.log; then cat /tmp/gpurun_last.log | tail -60; exit 0; fi
  sleep 120
done
echo "gave up after 40 transient answers"; tail -5 /tmp/gpurun_last.log
