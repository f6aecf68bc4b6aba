#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy / transient" (nothing charged)
log="$1"; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" "$log" || [ $rc -eq 3 ]; then
    echo "attempt $attempt: busy, retrying in 150 s" >> "${log}.retries"
    sleep 150
    continue
  fi
  break
done
exit $rc
