#!/bin/bash
# usage: [GPUS=n] tools/gpu_call.sh <log-name> <timeout-seconds> '<command>'   -- retries while the pod answers busy (exit 3 / transient)
log=gpurun_out/$1.log; shift
to=$1; shift
for i in $(seq 1 40); do
  if [ -n "$GPUS" ]; then gpurun --gpus $GPUS --timeout $to -- "$@" > $log 2>&1; else gpurun --timeout $to -- "$@" > $log 2>&1; fi
  rc=$?
  if grep -q "status=transient\|nothing was charged" $log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "done rc=$rc" >> $log
