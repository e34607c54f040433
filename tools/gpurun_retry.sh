#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> <timeout> <command...>   — retries while the pod answers busy / draining; log in gpurun_out/.<tag>.log
tag=$1; to=$2; shift 2
for k in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > gpurun_out/.$tag.log 2>&1
  if ! grep -q "status=transient\|rc=3\|no box" gpurun_out/.$tag.log; then break; fi
  sleep 150
done
echo done >> gpurun_out/.$tag.log
