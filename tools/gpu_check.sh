#!/bin/bash
# Short GPU session: GPU test suite, bench with and without programmatic dependent launch, LM launch list.
R=${1:-r1c}
mkdir -p gpurun_out
exec > gpurun_out/check_$R.log 2>&1
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 480 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 240 python bench.py --steps 20 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
TSLAM_PDL=0 timeout 120 python bench.py --steps 20 --no-extras 2>>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_nopdl_$R.json; cat gpurun_out/bench_nopdl_$R.json
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lm_$R.csv python tools/prof_lm.py 2 > gpurun_out/prof_lm.log 2>&1
tail -3 gpurun_out/bench_err_$R.log
ls -la gpurun_out
