"""Reduced-system solver alone on the tile pattern of the 500-keyframe global BA (16 leaves + 15 separators, two tiles each):
mean device time of the fused persistent kernel and of the wave kernels, and the task trace of one fused solve
(critical path = time from the first pop to the last done; per task type: count, mean wait, mean run)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import textslam_b200 as T
from chol_cases import NB, nd_pattern, spd_with_pattern, get_schedule, dev_chol_solve

ctx = T.Context(0)
rng = np.random.default_rng(1)
pat = nd_pattern(4); n = pat.shape[0] * NB
S = spd_with_pattern(rng, n, pat); b = rng.standard_normal(n)
ref = np.linalg.solve(S, b)
for mode, name in ((1, "fused"), (0, "waves")):
    x, ms, info, _ = dev_chol_solve(ctx, S, b, pat, mode, reps=20)
    print(f"{name}: {ms * 1e3:.1f} us per solve (tasks {info[0]}, waves {info[1]}, Tn {info[2]}), max err {np.abs(x - ref).max():.2e}")
x, ms, info, tr = dev_chol_solve(ctx, S, b, pat, 1, reps=1, want_trace=True)
tasks = get_schedule(n, pat)[0]
tr = tr.astype(np.int64)
t0 = tr[:, 0].min()
pop, ready, done, sm = tr[:, 0] - t0, tr[:, 1] - t0, tr[:, 2] - t0, tr[:, 3]
print(f"traced solve: {ms * 1e3:.1f} us by events, {done.max() / 1e3:.1f} us first pop -> last done, {len(set(sm.tolist()))} SMs used")
names = "FSUB"
for ty in range(4):
    m = tasks[:, 0] == ty
    if not m.any():
        continue
    rd = np.where(ready[m] > 0, ready[m], pop[m])
    print(f"  {names[ty]}: {m.sum():5d} tasks, wait {np.mean(rd - pop[m]) / 1e3:6.2f} us, run {np.mean(done[m] - rd) / 1e3:6.2f} us (max {np.max(done[m] - rd) / 1e3:.2f})")
# the F chain: every F task in queue order with its timeline
print("  F tasks (tile, pop, ready, done us):")
for t in np.nonzero(tasks[:, 0] == 0)[0]:
    print(f"    tile {tasks[t, 5]:3d}: {pop[t] / 1e3:7.2f} {ready[t] / 1e3:7.2f} {done[t] / 1e3:7.2f}")
fs = tr[tasks[:, 0] == 0][:, 4:15]
d = np.diff(fs, axis=1)
print("  F phases, cycles (mean over tasks): load", int(d[:, 0].mean()), "| per block column h: [T+U1 before potrf, potrf32_sym] =",
      [(int(d[:, 1 + 2 * h].mean()), int(d[:, 2 + 2 * h].mean())) for h in range(4)], "| tail", int((fs[:, 10] - fs[:, 9]).mean()))
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "chol_trace.npz"), trace=tr, tasks=tasks)
bt = np.nonzero(tasks[:, 0] == 3)[0]
print(f"  backward solve: first B ready at {pop[bt].min() / 1e3:.2f}, last done {done[bt].max() / 1e3:.2f}")
