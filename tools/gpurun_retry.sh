#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'   -- retries while gpurun answers "busy" (exit code 3)
for attempt in 1 2 3 4 5 6 7 8; do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpurun_retry] busy (attempt $attempt), sleeping 90 s" >&2
    sleep 90
done
exit 3
