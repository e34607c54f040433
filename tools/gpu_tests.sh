#!/bin/bash
# GPU test suite + smoke + one short bench line (no profiles)
R=${1:-t}
mkdir -p gpurun_out
exec > gpurun_out/tests_$R.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
