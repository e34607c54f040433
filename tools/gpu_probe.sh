#!/bin/bash
# Probe the GPU box for the reference's third-party toolchain (SURVEY 8c / VERDICT r1 next-3a) and record the state of the suite.
mkdir -p gpurun_out
exec > gpurun_out/probe_r2.log 2>&1
set -x
nvidia-smi -L
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
find / \( -name 'ceres*' -o -name 'Eigen' -o -name 'eigen3' -o -name 'opencv2' -o -name 'libopencv*' -o -name 'libceres*' -o -name 'glog' -o -name 'libglog*' -o -name 'sophus' \) -not -path '/proc/*' -not -path '/sys/*' 2>/dev/null | head -50
python -c "import pyceres" ; python -c "import cv2; print(cv2.__version__, cv2.__file__)"
ls baseline/_ref 2>&1 | head
dpkg -l | grep -iE "ceres|eigen|opencv|glog|suitesparse" | head
timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 200 python bench.py --steps 20 2>gpurun_out/bench_err_probe.log | tail -1 > gpurun_out/bench_probe.json; cat gpurun_out/bench_probe.json
