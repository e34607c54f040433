#!/bin/bash
# Cholesky class trace (CUDA events between launches, warm caches) with and without two-tile panels + ncu launch list.
R=${1:-p}
mkdir -p gpurun_out
exec > gpurun_out/prof_$R.log 2>&1
set -x
TSLAM_CHOL_TRACE=1 timeout 60 python tools/prof_lm.py 3 2>&1 | grep "tslam chol" | tail -2
TSLAM_CHOL_TRACE=1 TSLAM_CHOL_PAIR=0 timeout 60 python tools/prof_lm.py 3 2>&1 | grep "tslam chol" | tail -2
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lm_$R.csv python tools/prof_lm.py 2 > gpurun_out/prof_lm_$R.log 2>&1
