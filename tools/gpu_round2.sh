#!/bin/bash
# One GPU session of round 2: test suite, smoke, bench (both arms, with extras), then the profiles.
R=${1:-r2}
mkdir -p gpurun_out
exec > gpurun_out/round_$R.log 2>&1
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/bench_ref_$R.json; cat gpurun_out/bench_ref_$R.json
bash tools/gpu_prof2.sh $R
