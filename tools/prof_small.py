"""Where a small solve spends its time (C3 pose-only, C4 local BA): C-ABI internal laps (TSLAM_SETUP_TRACE), per-phase device
times of the LM loop, wall time per call with pageable and page-locked host buffers."""
import os, sys, time
os.environ["TSLAM_SETUP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import textslam_b200 as T
from textslam_b200 import synth
from bench import pinned_copy

ctx = T.Context(0)
for name, prob in (("c3", synth.c3_pose_only(seed=0)), ("c4", synth.c4_local_ba(seed=0))):
    print("=====", name, "pobs", prob.n_pobs, "tobs", prob.n_tobs, "cams", len(prob.cams), flush=True)
    for _ in range(3):
        s, _, _ = ctx.solve(prob.copy(), 10, want_trace=False)
    sys.stderr.flush()
    print("summary", s, flush=True)
    for kind, mk in (("pageable", lambda: prob.copy()), ("pinned", lambda: pinned_copy(prob, torch))):
        cps = [mk() for _ in range(20)]
        t0 = time.perf_counter()
        for c in cps:
            s, _, _ = ctx.solve(c, 10, want_trace=False)
        dt = (time.perf_counter() - t0) / 20
        print(f"{name} {kind}: wall {1e3*dt:.3f} ms/solve, C-ABI total_ms {s['total_ms']:.3f} setup_ms {s['setup_ms']:.3f} solve_ms {s['solve_ms']:.3f} its {s['iterations']}", flush=True)
    os.environ["TSLAM_SMALL_PROF"] = "1"
    ctx.solve(prob.copy(), 10, want_trace=False)
    sys.stderr.flush()
    os.environ.pop("TSLAM_SMALL_PROF")
    dev = ctx.upload(prob)
    for _ in range(3):
        ph, s = dev.lm_iterations(10)
    print("phases/iter", [round(float(x), 4) for x in ph], "its", s["iterations"], flush=True)
    dev.free()
