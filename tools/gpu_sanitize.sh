#!/bin/bash
# compute-sanitizer memcheck over the kernels written or changed late in round 2 (small cases)
mkdir -p gpurun_out
exec > gpurun_out/sanitize.log 2>&1
S="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
timeout 500 $S python -m pytest tests/test_gpu_small.py -m gpu -q -x -k "one_observation or c3_pose_only or landmarks_only or huber" 2>&1 | tail -6
timeout 300 $S python -m pytest tests/test_gpu_search3d.py -m gpu -q -x -k "edge or local_track" 2>&1 | tail -6
timeout 500 $S python -m pytest tests/test_gpu_orb.py -m gpu -q -x -k "row_stride or operator_call" 2>&1 | tail -6
timeout 400 $S python -m pytest tests/test_gpu_solve.py -m gpu -q -x -k "text_on and not central" 2>&1 | tail -6
