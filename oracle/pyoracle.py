"""ctypes binding of the CPU oracle (oracle/libtslam_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (textslam_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess
import numpy as np
from textslam_b200._abi import (BAProblemC, SolveOptionsC, SolveSummaryC, PT_NCOLS, TX_NCOLS, TRACE_COLS,
                                solve_options, c_dp, KP_DTYPE)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libtslam_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.tso_solve.restype = C.c_int
        _LIB.tso_eval_points.restype = C.c_int
        _LIB.tso_eval_text.restype = C.c_int
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def eval_points(prob, kind, want_J=True, n_threads=1):
    n, nc = prob.n_pobs, PT_NCOLS[kind]
    r = np.zeros((n, 2)); J = np.zeros((n, 2, nc)) if want_J else None
    pc = prob.as_c()
    lib().tso_eval_points(C.c_int(kind), C.byref(pc), _dp(r), _dp(J) if want_J else None, C.c_int(n_threads))
    return r, J


def eval_text(prob, kind, jac_mode, want_J=True, n_threads=1):
    n, nc = prob.n_tobs, TX_NCOLS[kind]
    r = np.zeros((n, 8)); J = np.zeros((n, 8, nc)) if want_J else None
    pc = prob.as_c()
    lib().tso_eval_text(C.c_int(kind), C.c_int(jac_mode), C.byref(pc), _dp(r), _dp(J) if want_J else None, C.c_int(n_threads))
    return r, J


def solve(prob, max_iters=10, text_jac_mode=0, n_threads=1, dense_full=0, want_trace=True, **kw):
    """Runs the Ceres-faithful LM loop; updates prob parameters in place. Returns (summary dict, final_residuals, trace)."""
    opt = solve_options(max_iters, text_jac_mode, n_threads, dense_full, **kw)
    summ = SolveSummaryC()
    fr = np.zeros(2 * prob.n_pobs + 8 * prob.n_tobs)
    tr = np.full((max_iters + 1, TRACE_COLS), np.nan)
    pc = prob.as_c()
    rc = lib().tso_solve(C.byref(pc), C.byref(opt), C.byref(summ), _dp(fr), _dp(tr) if want_trace else None)
    assert rc == 0
    return summ.as_dict(), fr, tr


def point_ambient(cam, host, rho, ray_xy, uv, K4, w):
    r = np.zeros(2); J = np.zeros((2, 15))
    f = lambda a: _dp(np.ascontiguousarray(a, dtype=np.float64))
    lib().tso_point_ambient(f(cam), f(host), C.c_double(rho), f(ray_xy), f(uv), f(K4), f(w), _dp(r), _dp(J))
    return r, J


def quat_plus(x, d):
    out = np.zeros(4)
    f = lambda a: _dp(np.ascontiguousarray(a, dtype=np.float64))
    lib().tso_quat_plus(f(x), f(d), _dp(out))
    return out
