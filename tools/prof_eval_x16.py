"""ncu driver: the residual+Jacobian kernel on the C5 x16 replica (1.6M evaluations, 429 MB per launch > L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import textslam_b200 as T
from textslam_b200 import synth
ctx = T.Context(0)
big = synth.make_ba_problem(seed=1, n_kf=500, n_lm=400000, obs_per_lm=4, band=10, fixed_cams=(0, 1), w_point=1.0, perturb=False)
d = ctx.upload(big)
ms = d.eval_points(T.PT_BA_NW, reps=5, flush_l2=True)
print("x16 eval ms/launch", ms, "GB/s", 268 * big.n_pobs / ms / 1e6)
