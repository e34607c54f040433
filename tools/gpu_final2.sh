#!/bin/bash
# Final GPU session of round 2: whole test suite, smoke, both bench arms, launch lists, ncu captures of the two new persistent kernels.
R=${1:-r2_final}
mkdir -p gpurun_out
exec > gpurun_out/final_$R.log 2>&1
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py --steps 20 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_$R.json; head -c 400 gpurun_out/bench_$R.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_$R.json; head -c 400 gpurun_out/bench_ref_$R.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lm_$R.csv python tools/prof_lm.py 2 > gpurun_out/prof_lm_$R.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_orb_$R.csv python tools/prof_orb.py > gpurun_out/prof_orb_$R.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_small_$R.csv python tools/prof_small.py > gpurun_out/prof_small_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:chol_fused_kernel -s 3 -c 1 -o gpurun_out/ncu_chol_fused_$R -f python tools/prof_lm.py 2 > gpurun_out/ncu_chol_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ba_small_kernel -s 4 -c 1 -o gpurun_out/ncu_ba_small_$R -f python tools/prof_small.py > gpurun_out/ncu_small_$R.log 2>&1
TSLAM_SMALL_PROF=1 timeout 100 python tools/prof_small.py 2>&1 | grep -v "^\[tslam" > gpurun_out/prof_small_phases_$R.log
ls -la gpurun_out | tail -14
