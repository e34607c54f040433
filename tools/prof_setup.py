"""Host-side cost split of one tslam_solve call on C5 (TSLAM_SETUP_TRACE=1 prints the laps to stderr)."""
import os, sys, time
os.environ["TSLAM_SETUP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import textslam_b200 as T
from textslam_b200 import synth

ctx = T.Context(0)
base = synth.c5_global_ba()
for it in range(3):
    prob = base.copy()
    t0 = time.perf_counter()
    s, _, _ = T.Optimizer(ctx).GlobalBA(prob)
    print(f"call {it}: wall {1e3 * (time.perf_counter() - t0):.2f} ms  its {s['iterations']} setup {s['setup_ms']:.2f} solve {s['solve_ms']:.2f} total {s['total_ms']:.2f}", file=sys.stderr)
