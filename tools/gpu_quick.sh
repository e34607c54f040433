#!/bin/bash
# Quick GPU iteration: GPU test suite + one bench line (no side measurements).
R=${1:-q}
mkdir -p gpurun_out
exec > gpurun_out/quick_$R.log 2>&1
set -x
timeout 480 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 200 python bench.py --steps 20 --no-extras 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
tail -5 gpurun_out/bench_err_$R.log
