#!/bin/bash
# A/B of one environment switch: GPU test suite with the default, then the bench line with and without the switch ($2=VAR=VALUE).
R=${1:-ab}
mkdir -p gpurun_out
exec > gpurun_out/ab_$R.log 2>&1
set -x
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 150 python bench.py --steps 20 --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_${R}_default.json
env $2 timeout 150 python bench.py --steps 20 --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_${R}_switch.json
