#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the shared-memory-heavy kernels of round 2
mkdir -p gpurun_out
exec > gpurun_out/racecheck.log 2>&1
S="compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 8"
timeout 600 $S python -m pytest tests/test_gpu_small.py -m gpu -q -x -k "c4_local_ba or landmarks_only" 2>&1 | tail -8
timeout 600 $S python -m pytest tests/test_gpu_orb.py -m gpu -q -x -k "operator_call" 2>&1 | tail -12
timeout 900 $S python -m pytest tests/test_gpu_chol.py -m gpu -q -x -k "pattern_from_nonzeros" 2>&1 | tail -12
