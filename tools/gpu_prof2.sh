#!/bin/bash
# Round-2 profiles: launch lists (LM iterations, ORB) and ncu --set full captures of the kernels the bench line quotes.
R=${1:-r2}
mkdir -p gpurun_out
exec > gpurun_out/prof_$R.log 2>&1
set -x
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lm_$R.csv python tools/prof_lm.py 2 > gpurun_out/prof_lm_$R.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_orb_$R.csv python tools/prof_orb.py 64 > gpurun_out/prof_orb_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:point_eval_kernel -s 0 -c 2 -o gpurun_out/ncu_point_eval_x16_$R -f python tools/prof_eval_x16.py > gpurun_out/ncu_eval_x16_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:text_eval_kernel -s 0 -c 2 -o gpurun_out/ncu_text_eval_$R -f python tools/prof_text.py > gpurun_out/ncu_text_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:chol_fused_kernel -s 3 -c 1 -o gpurun_out/ncu_chol_fused_$R -f python tools/prof_lm.py 2 > gpurun_out/ncu_chol_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:schur_merged_kernel -s 2 -c 1 -o gpurun_out/ncu_schur_merged_$R -f python tools/prof_lm.py 2 > gpurun_out/ncu_schur_$R.log 2>&1
ls -la gpurun_out | tail -12
