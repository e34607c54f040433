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


def dense_ceres_lm_ba(prob, max_iters):
    """The same loop for a small bundle adjustment: free cameras (6 tangent columns each, applied through Plus) and free inverse depths
    (plain additive), one dense Jacobian, no Schur complement. The C++ oracle eliminates the landmarks first (an exact
    reformulation, Appendix A.7), so its trace must still agree."""
    K, w, delta = prob.K_point, prob.w_point, prob.huber_point
    cams, rho = prob.cams.copy(), prob.rho.copy()
    fc = [int(c) for c in np.nonzero(prob.cam_fixed == 0)[0]]
    fl = [int(l) for l in np.nonzero(prob.rho_fixed == 0)[0]]
    n = 6 * len(fc) + len(fl)

    def residuals(cams_, rho_):
        return np.array([point_residual(cams_[prob.p_cam[i]], cams_[prob.p_host[i]], rho_[prob.p_lm[i]], prob.p_ray[i], prob.p_uv[i], K, w) for i in range(prob.n_pobs)])

    def apply(cams_, rho_, step):
        c2, r2 = cams_.astype(step.dtype if np.iscomplexobj(step) else float), rho_.astype(step.dtype if np.iscomplexobj(step) else float)
        for k, c in enumerate(fc):
            c2[c] = plus(c2[c], step[6 * k:6 * k + 6])
        for k, l in enumerate(fl):
            r2[l] = r2[l] + step[6 * len(fc) + k]
        return c2, r2

    def jacobian(cams_, rho_, h=1e-30):
        J = np.zeros((prob.n_pobs, 2, n))
        for k in range(n):
            d = np.zeros(n, dtype=complex); d[k] = 1j * h
            J[:, :, k] = residuals(*apply(cams_, rho_, d)).imag / h
        return J

    def robustify(r):
        s = (r ** 2).sum(1)
        out = s > delta * delta
        rho_l = np.where(out, 2 * delta * np.sqrt(np.maximum(s, 1e-300)) - delta * delta, s)
        return 0.5 * rho_l.sum(), np.where(out, np.sqrt(delta / np.sqrt(np.maximum(s, 1e-300))), 1.0)

    def xnorm(cams_, rho_):
        return np.sqrt(sum((cams_[c] ** 2).sum() for c in fc) + sum(rho_[l] ** 2 for l in fl))

    r = residuals(cams, rho); cost, sq = robustify(r)
    Jm = (jacobian(cams, rho) * sq[:, None, None]).reshape(-1, n); rc = (r * sq[:, None]).reshape(-1)
    scale = 1.0 / (1.0 + np.sqrt((Jm ** 2).sum(0)))
    radius, decrease_factor = 1e4, 2.0
    trace = [(cost, radius, 0.0, 1)]
    for _ in range(max_iters):
        Js = Jm * scale
        H, g = Js.T @ Js, Js.T @ rc
        step = np.linalg.solve(H + np.diag(np.clip(np.diag(H), 1e-6, 1e32) / radius), -g) * scale
        Jd = Jm @ step
        model_change = -(Jd @ (rc + 0.5 * Jd))
        c_new, r_new_p = apply(cams, rho, step)
        res_new = residuals(c_new, r_new_p); cost_new, sq_new = robustify(res_new)
        if not model_change > 0:
            radius /= decrease_factor; decrease_factor *= 2; trace.append((cost, radius, 0.0, -1)); continue
        step_norm = np.sqrt(sum(((c_new[c] - cams[c]) ** 2).sum() for c in fc) + sum((r_new_p[l] - rho[l]) ** 2 for l in fl))
        if step_norm <= 1e-8 * (xnorm(cams, rho) + 1e-8) or abs(cost - cost_new) <= 1e-6 * cost:
            trace.append((cost, radius, 0.0, 0)); break
        rel = (cost - cost_new) / model_change
        if rel > 1e-3:
            cams, rho, r, cost, sq = c_new, r_new_p, res_new, cost_new, sq_new
            Jm = (jacobian(cams, rho) * sq[:, None, None]).reshape(-1, n); rc = (r * sq[:, None]).reshape(-1)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3)); decrease_factor = 2.0
            trace.append((cost, radius, rel, 1))
        else:
            radius /= decrease_factor; decrease_factor *= 2; trace.append((cost_new, radius, rel, 0))
    return np.array(trace), cams, rho


@pytest.mark.parametrize("seed", [12, 13])
def test_small_ba_trace_matches_second_reading(oracle, seed):
    prob = synth.make_ba_problem(seed=seed, n_kf=4, n_lm=24, obs_per_lm=3, band=4, fixed_cams=(0, 1), rot_noise=3e-2, trans_noise=3e-2, rho_noise=0.1)
    tr_py, cams_py, rho_py = dense_ceres_lm_ba(prob, 10)
    a = prob.copy()
    _, _, tr = oracle.solve(a, 10)
    tr = tr[~np.isnan(tr[:, 0])]
    assert len(tr) == len(tr_py) and np.array_equal(tr[:, 3], tr_py[:, 3])
    assert np.allclose(tr[:, 0], tr_py[:, 0], rtol=1e-7) and np.allclose(tr[:, 1], tr_py[:, 1], rtol=1e-4)
    assert np.abs(a.cams - cams_py).max() < 1e-7 and np.abs(a.rho - rho_py).max() < 1e-7
