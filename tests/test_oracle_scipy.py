"""Independent check of the oracle's optimum (not of its iterates): scipy.optimize.least_squares with a Huber loss on the same
robustified objective must reach the cost and the parameters the Ceres-style LM loop of oracle/ba_lm.cpp converges to.
Ceres applies HuberLoss to the squared norm of a residual BLOCK (s = |r|^2: rho(s) = s for s <= d^2, 2 d sqrt(s) - d^2 beyond);
scipy applies its loss per scalar residual, so each block is handed over as its norm with f_scale = d — then
f_scale^2 * rho_scipy(|r|^2 / f_scale^2) is exactly Ceres' rho(s). The camera is parametrised through the same quaternion Plus."""
import numpy as np
import pytest
from scipy.optimize import least_squares
from textslam_b200 import synth
from textslam_b200._abi import PT_BA


def _block_norms(oracle, prob):
    r, _ = oracle.eval_points(prob, PT_BA, want_J=False)
    return np.sqrt((r ** 2).sum(1))


def _robust_cost(norms, delta):
    s = norms ** 2
    rho = np.where(s <= delta * delta, s, 2.0 * delta * np.sqrt(s) - delta * delta)
    return 0.5 * rho.sum()


def test_pose_only_optimum_matches_scipy(oracle):
    prob = synth.c3_pose_only(seed=5, n_pobs=300, n_planes=0)
    delta = prob.huber_point
    assert delta > 0 and prob.cam_fixed[0] == 0 and prob.cam_fixed[1:].all() and prob.rho_fixed.all()
    q0, t0 = prob.cams[0, :4].copy(), prob.cams[0, 4:].copy()

    def fun(x):
        p = prob.copy()
        p.cams[0, :4] = oracle.quat_plus(q0, x[:3]); p.cams[0, 4:] = t0 + x[3:]
        return _block_norms(oracle, p)

    sol = least_squares(fun, np.zeros(6), loss="huber", f_scale=delta, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-12, x_scale=1e-2)
    cost_scipy = _robust_cost(fun(sol.x), delta)
    a = prob.copy()
    summ, _, _ = oracle.solve(a, 60, function_tolerance=1e-16, parameter_tolerance=1e-14, gradient_tolerance=1e-14)
    assert abs(summ["final_cost"] - cost_scipy) <= 1e-9 * cost_scipy
    assert abs(summ["final_cost"] - _robust_cost(_block_norms(oracle, a), delta)) <= 1e-12 * cost_scipy   # the summary reports the robustified cost
    q_s = oracle.quat_plus(q0, sol.x[:3])
    assert np.abs(a.cams[0, :4] - q_s).max() < 1e-6 and np.abs(a.cams[0, 4:] - (t0 + sol.x[3:])).max() < 1e-6
    assert (_block_norms(oracle, a) > delta).sum() >= 5   # the planted outliers sit on the linear branch of the loss


def test_small_ba_optimum_is_a_stationary_point_for_scipy(oracle):
    """Two free keyframes + 20 free inverse depths (Schur-eliminated in the oracle, plain dense unknowns for scipy): restarted from the
    oracle's solution, scipy's trust-region solver finds nothing better (the point is a local minimum of the same robustified objective);
    started from the initial estimate it does not end below it either."""
    prob = synth.make_ba_problem(seed=9, n_kf=4, n_lm=20, obs_per_lm=3, band=4, fixed_cams=(0, 1))
    delta = prob.huber_point
    free = np.nonzero(prob.cam_fixed == 0)[0]
    a = prob.copy()
    summ, _, _ = oracle.solve(a, 100, function_tolerance=1e-16, parameter_tolerance=1e-14, gradient_tolerance=1e-14)
    cost_oracle = summ["final_cost"]
    assert abs(cost_oracle - _robust_cost(_block_norms(oracle, a), delta)) <= 1e-12 * cost_oracle

    def make_fun(base):
        q0, t0, rho0 = base.cams[free, :4].copy(), base.cams[free, 4:].copy(), base.rho.copy()

        def fun(x):
            p = base.copy()
            for k, c in enumerate(free):
                p.cams[c, :4] = oracle.quat_plus(q0[k], x[6 * k:6 * k + 3]); p.cams[c, 4:] = t0[k] + x[6 * k + 3:6 * k + 6]
            p.rho[:] = rho0 + x[6 * len(free):]
            return _block_norms(oracle, p)
        return fun

    n = 6 * len(free) + len(prob.rho)
    restart = least_squares(make_fun(a), np.zeros(n), loss="huber", f_scale=delta, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-12, x_scale=1e-2, max_nfev=30)
    cost_restart = _robust_cost(make_fun(a)(restart.x), delta)
    assert cost_restart >= cost_oracle * (1 - 1e-8)          # nothing better in the neighbourhood
    assert np.abs(restart.x).max() < 1e-4                    # and scipy does not walk away from it
    cold = least_squares(make_fun(prob), np.zeros(n), loss="huber", f_scale=delta, method="trf", x_scale=1e-2, max_nfev=40)
    assert _robust_cost(make_fun(prob)(cold.x), delta) >= cost_oracle * (1 - 1e-6)
