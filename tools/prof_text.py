"""ncu driver: the text residual+Jacobian kernel (nume_BAText, analytic mode) on the text-on C5 problem (25k blocks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import textslam_b200 as T
from textslam_b200 import synth
ctx = T.Context(0)
prob = synth.c5_global_ba(seed=0, n_planes=1000)
d = ctx.upload(prob)
ms = d.eval_text(T.TX_BA, T.JAC_ANALYTIC, reps=5, flush_l2=True)
print("text eval ms/launch", ms, "GB/s", 1280 * prob.n_tobs / ms / 1e6)
