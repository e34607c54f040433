#!/bin/bash
# One GPU session: full GPU test suite, bench (both arms), ncu launch lists and full captures of the top kernels.
R=${1:-r1}
mkdir -p gpurun_out
exec > gpurun_out/round_$R.log 2>&1
set -x
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/bench_ref_$R.json; cat gpurun_out/bench_ref_$R.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lm_$R.csv python tools/prof_lm.py 2 > gpurun_out/prof_lm.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_orb_$R.csv python tools/prof_orb.py 64 > gpurun_out/prof_orb.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:point_eval_kernel -s 0 -c 2 -o gpurun_out/prof_point_eval_x16_$R -f python tools/prof_eval_x16.py > gpurun_out/prof_eval_x16.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:potrf_trsm_kernel -s 20 -c 1 -o gpurun_out/prof_potrf_trsm_$R -f python tools/prof_lm.py 2 > gpurun_out/prof_potrf.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:schur_block_kernel -s 4 -c 2 -o gpurun_out/prof_schur_block_$R -f python tools/prof_lm.py 2 > gpurun_out/prof_schur.log 2>&1
timeout 100 python tools/prof_setup.py 2> gpurun_out/prof_setup_$R.log
ls -la gpurun_out
