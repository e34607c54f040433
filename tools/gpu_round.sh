#!/bin/bash
# One GPU session: full GPU test suite, bench (both arms), ncu launch lists and full captures of the top kernels.
R=${1:-r1}
mkdir -p gpurun_out
exec > gpurun_out/round_$R.log 2>&1
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json
python bench.py --impl reference --steps 3 --warmup 1 | tail -1 > gpurun_out/bench_ref_$R.json; cat gpurun_out/bench_ref_$R.json
python tools/prof_orb.py 64
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lm_$R.csv python tools/prof_lm.py 2 > gpurun_out/prof_lm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_orb_$R.csv python tools/prof_orb.py 64 > gpurun_out/prof_orb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:point_eval_kernel -s 0 -c 2 -o gpurun_out/prof_point_eval_x16_$R -f python tools/prof_eval_x16.py > gpurun_out/prof_eval_x16.log 2>&1
python tools/prof_eval_x16.py 2>&1 | tail -1; TSLAM_EVAL_OCC8=1 python tools/prof_eval_x16.py 2>&1 | tail -1
ls -la gpurun_out
