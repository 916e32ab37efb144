#!/bin/bash
# gpurun with retries while the pod answers "busy" (nothing is charged for those); usage: gpurun_retry.sh OUTFILE [gpurun args...]
out=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if ! grep -q "status=transient" "$out"; then exit 0; fi
  sleep 60
done
exit 3
