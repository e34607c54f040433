"""Second reading of the Levenberg-Marquardt loop of ceres::Solve as ~60 lines of dense numpy —
its own functor (tests/test_oracle_functors_py.py), its own complex-step Jacobian through the quaternion Plus, its own Huber
corrector, Jacobi scaling, damped normal equations, step test and radius update. oracle/ba_lm.cpp (Schur-eliminated, analytic
derivatives, C++) must produce the same trace: accept / reject sequence, cost and trust-region radius of every iteration."""
import numpy as np
import pytest
from textslam_b200 import synth
from test_oracle_functors_py import point_residual, plus


def dense_ceres_lm(blocks, x0_cams, free_cam, delta, max_iters):
    """blocks: list of (cam_idx, host_idx, rho, ray, uv); only camera `free_cam` is optimised. Returns trace rows
    (cost, radius, relative_decrease, accepted) like the oracle's, iteration 0 first."""
    cams = x0_cams.copy()
    K, w = blocks["K"], blocks["w"]

    def residuals(c):
        cc = cams.copy(); cc[free_cam] = c
        return np.array([point_residual(cc[a], cc[h], rho, ray, uv, K, w) for a, h, rho, ray, uv in blocks["obs"]])

    def robustify(r):
        s = (r ** 2).sum(1)
        out = s > delta * delta
        rho = np.where(out, 2 * delta * np.sqrt(np.maximum(s, 1e-300)) - delta * delta, s)
        sq = np.where(out, np.sqrt(delta / np.sqrt(np.maximum(s, 1e-300))), 1.0)      # sqrt(rho')
        return 0.5 * rho.sum(), sq

    def jacobian(c, h=1e-30):
        # complex-step derivative (the functor and Plus are analytic): exact to rounding even where a point sits close to the camera plane
        J = np.zeros((len(blocks["obs"]), 2, 6))
        for k in range(6):
            d = np.zeros(6, dtype=complex); d[k] = 1j * h
            cc = cams.astype(complex); cc[free_cam] = plus(c.astype(complex), d)
            J[:, :, k] = np.array([point_residual(cc[a], cc[hh], rho, ray, uv, K, w) for a, hh, rho, ray, uv in blocks["obs"]]).imag / h
        return J

    x = cams[free_cam].copy()
    r = residuals(x); cost, sq = robustify(r)
    J = jacobian(x) * sq[:, None, None]; rc = (r * sq[:, None]).reshape(-1); Jm = J.reshape(-1, 6)
    scale = 1.0 / (1.0 + np.sqrt((Jm ** 2).sum(0)))                                    # jacobi_scaling, once
    radius, decrease_factor = 1e4, 2.0
    trace = [(cost, radius, 0.0, 1)]
    for _ in range(max_iters):
        Js = Jm * scale
        g = Js.T @ rc
        H = Js.T @ Js
        D2 = np.clip(np.diag(H), 1e-6, 1e32) / radius
        d_scaled = np.linalg.solve(H + np.diag(D2), -g)
        step = d_scaled * scale
        Jd = Jm @ step
        model_change = -(Jd @ (rc + 0.5 * Jd))
        x_new = plus(x, step)
        r_new = residuals(x_new); cost_new, sq_new = robustify(r_new)
        if not model_change > 0:
            radius /= decrease_factor; decrease_factor *= 2
            trace.append((cost, radius, 0.0, -1)); continue
        step_norm = np.linalg.norm(x_new - x)
        if step_norm <= 1e-8 * (np.linalg.norm(x) + 1e-8):
            trace.append((cost, radius, 0.0, 0)); break
        if abs(cost - cost_new) <= 1e-6 * cost:
            trace.append((cost, radius, 0.0, 0)); break
        rel = (cost - cost_new) / model_change
        if rel > 1e-3:
            x, r, cost, sq = x_new, r_new, cost_new, sq_new
            J = jacobian(x) * sq[:, None, None]; rc = (r * sq[:, None]).reshape(-1); Jm = J.reshape(-1, 6)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3)); decrease_factor = 2.0
            trace.append((cost, radius, rel, 1))
        else:
            radius /= decrease_factor; decrease_factor *= 2
            trace.append((cost_new, radius, rel, 0))
    return np.array(trace), x


@pytest.mark.parametrize("seed,rot_noise", [(3, 1e-2), (4, 8e-2), (6, 0.25), (7, 0.6)])
def test_pose_only_trace_matches_second_reading(oracle, seed, rot_noise):
    prob = synth.make_ba_problem(seed=seed, n_kf=1, n_lm=150, obs_per_lm=1, band=20, fixed_cams=(), n_ext=20, frac_ext_lm=1.0,
                                 rot_noise=rot_noise, trans_noise=rot_noise)
    blocks = {"K": prob.K_point, "w": prob.w_point,
              "obs": [(int(prob.p_cam[i]), int(prob.p_host[i]), float(prob.rho[prob.p_lm[i]]), prob.p_ray[i], prob.p_uv[i]) for i in range(prob.n_pobs)]}
    tr_py, x_py = dense_ceres_lm(blocks, prob.cams, 0, prob.huber_point, 10)
    a = prob.copy()
    summ, _, tr = oracle.solve(a, 10)
    tr = tr[~np.isnan(tr[:, 0])]
    assert len(tr) == len(tr_py), (tr, tr_py)
    assert np.array_equal(tr[:, 3], tr_py[:, 3])                                      # accepted / rejected / invalid / terminated
    assert np.allclose(tr[:, 0], tr_py[:, 0], rtol=1e-7)                              # cost after every iteration
    assert np.allclose(tr[:, 1], tr_py[:, 1], rtol=1e-4)                              # trust-region radius
    assert np.abs(a.cams[0] - x_py).max() < 1e-7
    if rot_noise >= 0.5:   # the hard start ends with three rejected steps in a row: radius / 2, / 4, / 8 (decrease_factor doubling)
        assert tr[-3:, 3].tolist() == [0, 0, 0] and len(tr) == 11
        assert np.allclose(tr[-3:, 1] / tr[-4:-1, 1], [0.5, 0.25, 0.125], rtol=1e-12)
