#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <tag> <timeout> <command...>
g=$1; tag=$2; to=$3; shift 3
for k in $(seq 1 8); do
  /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$to" -- "$@" > gpurun_out/.$tag.log 2>&1
  if ! grep -q "status=transient\|rc=3\|no box\|busy" gpurun_out/.$tag.log; then break; fi
  sleep 200
done
echo done >> gpurun_out/.$tag.log
