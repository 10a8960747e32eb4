#!/usr/bin/env bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 / "transient"): nothing is charged
# for a refused call.  Usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    rc=$?
    if [ $rc -ne 3 ] && ! grep -q "status=transient" <<<"$out"; then
        echo "$out"
        exit $rc
    fi
    echo "[retry $attempt] no GPU slot; sleeping 120 s" >&2
    sleep 120
done
echo "$out"
exit 3
