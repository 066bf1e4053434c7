#!/bin/bash
# usage: [GPURUN_FLAGS="--gpus 2"] gpurun_retry.sh <timeout_s> <max_tries> <command...>   -- retries while the pod answers busy/transient (exit 3)
T=$1; N=$2; shift 2
for i in $(seq 1 $N); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout $T -- "$@" > /tmp/gpurun_retry.log 2>&1; rc=$?
  if grep -q "status=transient\|status=busy" /tmp/gpurun_retry.log || [ $rc -eq 3 ]; then sleep 200; continue; fi
  break
done
tail -40 /tmp/gpurun_retry.log; echo "gpurun rc=$rc tries=$i"
