#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  usage: gpurun_retry.sh TAG TIMEOUT 'command' [extra gpurun flags]
tag=$1; to=$2; cmd=$3; shift 3
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" --timeout "$to" -- "$cmd" > gpurun_out/call_${tag}.log 2>&1
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 90
done
exit 3
