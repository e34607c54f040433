#!/bin/bash
# One full GPU session: GPU test suite, smoke, bench (both arms), ncu launch lists and full captures of the top kernels.
R=${1:-r1}
mkdir -p gpurun_out
exec > gpurun_out/round_$R.log 2>&1
set -x
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 30 2>gpurun_out/bench_err_$R.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/bench_ref_$R.json; cat gpurun_out/bench_ref_$R.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lm_$R.csv python tools/prof_lm.py 2 > gpurun_out/prof_lm_$R.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_orb_$R.csv python tools/prof_orb.py 64 > gpurun_out/prof_orb_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:point_eval_kernel -s 0 -c 2 -o gpurun_out/prof_point_eval_x16_$R -f python tools/prof_eval_x16.py > gpurun_out/prof_eval_x16_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:potrf_trsm_kernel -s 12 -c 1 -o gpurun_out/prof_potrf_trsm_$R -f python tools/prof_lm.py 2 > gpurun_out/prof_potrf_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:schur_merged_kernel -s 2 -c 1 -o gpurun_out/prof_schur_merged_$R -f python tools/prof_lm.py 2 > gpurun_out/prof_schur_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:syrk_wave_kernel -s 10 -c 1 -o gpurun_out/prof_syrk_$R -f python tools/prof_lm.py 2 > gpurun_out/prof_syrk_$R.log 2>&1
TSLAM_SETUP_TRACE=1 timeout 100 python tools/prof_setup.py 2> gpurun_out/prof_setup_$R.log
TSLAM_CHOL_TRACE=1 timeout 60 python tools/prof_lm.py 3 2>&1 | grep "tslam chol" | tail -2
ls -la gpurun_out
