#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3, nothing charged).
#   tools/gpurun_retry.sh [--gpus N] --timeout S -- '<command>'
for attempt in $(seq 1 20); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 90
done
exit 3
