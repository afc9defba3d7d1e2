#!/bin/bash
# usage: scripts/gpu_retry.sh <out-file> <timeout-s> <command...>   - retries gpurun while the pod answers "transient" / busy
out=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > "$out" 2>&1
  rc=$?
  if grep -q "status=transient\|exit code 3\|answers busy" "$out" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpu_retry done rc=$rc attempts=$i" >> "$out"
