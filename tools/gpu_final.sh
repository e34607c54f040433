#!/bin/bash
# Final verification: GPU test suite, smoke, the full bench line.
R=${1:-f}
mkdir -p gpurun_out
exec > gpurun_out/final_$R.log 2>&1
set -x
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 30 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
