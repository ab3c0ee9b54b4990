#!/bin/bash
# Usage: tools/gpurun_retry.sh <logfile> <timeout_s> [--gpus N] -- <command>
# Retries `gpurun` while the pod answers "busy" (exit code 3: nothing charged), every 90 s, for up to ~60 min.
log=$1; to=$2; shift 2
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then
    echo "gpurun finished rc=$rc (attempt $attempt)" >> "$log"
    exit $rc
  fi
  sleep 90
done
echo "gpurun: gave up after 40 busy answers" >> "$log"
exit 3
