#!/bin/bash
# usage: tools/gpu_retry.sh <out-file> <timeout> [--gpus N] -- <command>   : retry gpurun while the pod answers "transient"/busy
OUT=$1; shift; TO=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" > $OUT 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" $OUT || grep -q "retry in a few minutes" $OUT; then sleep 90; continue; fi
  break
done
