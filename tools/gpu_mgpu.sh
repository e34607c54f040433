#!/bin/bash
# Two-GPU session: landmark-sharded solve == single-GPU solve (tools/mgpu_check.py), then the bench line at N=2.
R=${1:-m}
mkdir -p gpurun_out
exec > gpurun_out/mgpu_$R.log 2>&1
set -x
N=${2:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v Warning | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_n${N}_$R.json; cat gpurun_out/bench_n${N}_$R.json
tail -5 gpurun_out/bench_err_$R.log
