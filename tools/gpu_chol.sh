#!/bin/bash
# Fused Cholesky bring-up: unit tests of the solver alone, trace on the C5 tile pattern, then the whole suite + bench.
R=${1:-c1}
mkdir -p gpurun_out
exec > gpurun_out/chol_$R.log 2>&1
set -x
timeout 300 python -m pytest tests/test_gpu_chol.py -q -x 2>&1 | tail -15
timeout 120 python tools/prof_chol.py 2>&1 | tail -60
timeout 480 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 200 python bench.py --steps 20 --no-extras 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
