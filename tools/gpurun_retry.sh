#!/bin/bash
log=$1; shift; to=$1; shift; n=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient\|status=busy" $log; then break; fi
  sleep 60
done
