"""Small driver for ncu: C5 global BA, a few LM iterations on the device-resident problem (plus the
stand-alone residual+Jacobian kernel). Used for the launch lists / ncu captures under profiles/."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import textslam_b200 as T
from textslam_b200 import synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = T.Context(0)
prob = synth.c5_global_ba(seed=0)
dev = ctx.upload(prob)
phases, summ = dev.lm_iterations(iters)
phases, summ = dev.lm_iterations(iters)
print(json.dumps({"phases_ms_per_iter": phases, "summary": summ}))
ms = dev.eval_points(T.PT_BA_NW, reps=5, flush_l2=True)
print("eval_points C5 ms/launch (L2 flushed):", ms, "GB/s:", 268 * prob.n_pobs / ms / 1e6)
ms = dev.eval_points(T.PT_BA_NW, reps=5, flush_l2=False)
print("eval_points C5 ms/launch (L2 warm):", ms, "GB/s:", 268 * prob.n_pobs / ms / 1e6)
