"""Phase split of the small solves (BASELINE configs: C3 pose-only, C4 local BA), device-resident and end to end."""
import os, sys, time, json
os.environ["TSLAM_SETUP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import textslam_b200 as T
from textslam_b200 import synth

ctx = T.Context(0)
names = ['eval', 'prep', 'build', 'allreduce', 'chol', 'backsub', 'model', 'iter']
for name, prob in (("c3", synth.c3_pose_only(seed=0)), ("c4", synth.c4_local_ba(seed=0))):
    dev = ctx.upload(prob)
    dev.lm_iterations(10)
    ph, summ = dev.lm_iterations(10)
    print(name, {n: round(v * 1e3, 1) for n, v in zip(names, ph)}, "its", summ["iterations"], file=sys.stderr)
    l0 = T._lib.lib().tslam_launch_count()
    dev.lm_iterations(10)
    print(name, "launches per solve", T._lib.lib().tslam_launch_count() - l0, file=sys.stderr)
    dev.free()
    for k in range(3):
        q = prob.copy(); t0 = time.perf_counter(); s, _, _ = ctx.solve(q, 10, want_trace=False)
        print(name, f"e2e {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr)
