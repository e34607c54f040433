#!/bin/bash
# Fused Cholesky: solver-alone tests + trace (no full suite)
R=${1:-c2}
mkdir -p gpurun_out
exec > gpurun_out/chol_$R.log 2>&1
set -x
timeout 300 python -m pytest tests/test_gpu_chol.py -q -x 2>&1 | tail -5
timeout 120 python tools/prof_chol.py 2>&1 | tail -60
