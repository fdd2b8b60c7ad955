#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box or slot free: nothing is charged).
#   tools/gpurun_retry.sh <log> [gpurun options] -- '<command>'
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy" "$log" || grep -q "retry in a few minutes" "$log"; then
    echo "[retry $attempt] $(tail -2 "$log" | head -1 | cut -c1-160)" >&2
    sleep 150
    continue
  fi
  break
done
tail -40 "$log" | cut -c1-400
