#!/bin/bash
# usage: tools/gpurun_retry.sh <log> [gpurun args...] -- '<command>'   (retries while the pod answers busy/transient)
log=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" "$log" && ! grep -q "charged=[1-9]" "$log"; then sleep 150; continue; fi
  break
done
