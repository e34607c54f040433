#!/bin/bash
R=${1:-t1}
mkdir -p gpurun_out
exec > gpurun_out/tma_$R.log 2>&1
set -x
timeout 300 python -m pytest tests/test_gpu_eval.py tests/test_gpu_solve.py -q -x 2>&1 | tail -8
timeout 200 python - <<'PY'
import sys; sys.path.insert(0, '.')
import textslam_b200 as T
from textslam_b200 import synth
ctx = T.Context(0)
prob = synth.c5_global_ba(seed=0, n_planes=1000)
d = ctx.upload(prob)
for mode, name in ((T.JAC_ANALYTIC, "ldg taps"), (T.JAC_ANALYTIC_TMA, "tma staged")):
    d.eval_text(T.TX_BA, mode, reps=3, flush_l2=True)
    for fl in (True, False):
        ms = d.eval_text(T.TX_BA, mode, reps=20, flush_l2=fl)
        print(f"text eval {name:10s} flush_l2={fl}: {ms*1e3:.2f} us / launch, {1280*prob.n_tobs/ms/1e6:.0f} GB/s")
PY
