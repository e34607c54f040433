#!/bin/bash
# ORB: bit-exactness tests + batch timing with the closed-form and the bisection FAST measure.
R=${1:-o}
mkdir -p gpurun_out
exec > gpurun_out/orb_$R.log 2>&1
set -x
timeout 300 python -m pytest tests/test_gpu_orb.py tests/test_gpu_orb_stages.py -m gpu -q -x 2>&1 | tail -5
timeout 100 python tools/prof_orb.py 64
TSLAM_FAST_BISECT=1 timeout 100 python tools/prof_orb.py 64
