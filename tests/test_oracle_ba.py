"""CPU tests of the oracle itself (no GPU): Jacobian self-consistency (dual numbers vs central
differences vs closed form), Ceres Plus semantics, LM convergence to ground truth, Schur == full solve."""
import numpy as np
import pytest
from textslam_b200 import synth
from textslam_b200._abi import PT_BA, PT_BA_NW, PT_POSE, PT_RHO, TX_BA, TX_POSE, TX_THETA, JAC_ANALYTIC, JAC_CENTRAL_DIFF


def _fd_tangent(po, prob, i, eps=1e-6):
    """Central differences of the point functor through Ceres' Plus (independent of the Jets)."""
    cam, host = prob.cams[prob.p_cam[i]].copy(), prob.cams[prob.p_host[i]].copy()
    rho = prob.rho[prob.p_lm[i]]

    def f(dc, dh, drho):
        c = cam.copy(); h = host.copy()
        c[:4] = po.quat_plus(cam[:4], dc[:3]); c[4:] += dc[3:]
        h[:4] = po.quat_plus(host[:4], dh[:3]); h[4:] += dh[3:]
        r, _ = po.point_ambient(c, h, rho + drho, prob.p_ray[i], prob.p_uv[i], prob.K_point, prob.w_point)
        return r

    J = np.zeros((2, 13))
    for k in range(13):
        d = np.zeros(13); d[k] = eps
        J[:, k] = (f(d[:6], d[6:12], d[12]) - f(-d[:6], -d[6:12], -d[12])) / (2 * eps)
    return J


def test_point_jets_match_central_differences(oracle):
    prob = synth.c4_local_ba(seed=3, n_planes=0)
    r, J = oracle.eval_points(prob, PT_BA)
    for i in range(0, prob.n_pobs, 311):
        Jfd = _fd_tangent(oracle, prob, i)
        assert np.allclose(J[i], Jfd, rtol=2e-6, atol=2e-5), i


def test_point_kinds_are_column_subsets(oracle):
    prob = synth.c4_local_ba(seed=4, n_planes=0)
    r, J = oracle.eval_points(prob, PT_BA)
    rp, Jp = oracle.eval_points(prob, PT_POSE)
    assert np.array_equal(r, rp) and np.array_equal(J[:, :, :6], Jp)
    rn, Jn = oracle.eval_points(prob, PT_BA_NW)
    w = prob.w_point[0]
    assert np.allclose(rn * w, r, rtol=1e-14) and np.allclose(Jn * w, J, rtol=1e-12, atol=1e-300)
    rr, Jr = oracle.eval_points(prob, PT_RHO)
    assert np.allclose(rr, rn, rtol=1e-15) and np.allclose(Jr[:, :, 0], Jn[:, :, 12], rtol=1e-13)


def test_text_numeric_vs_analytic(oracle):
    prob = synth.c4_local_ba(seed=5, n_lm=50)
    rn, Jn = oracle.eval_text(prob, TX_BA, JAC_CENTRAL_DIFF)
    ra, Ja = oracle.eval_text(prob, TX_BA, JAC_ANALYTIC)
    assert np.array_equal(rn, ra)
    # the two agree except where the +-h stencil straddles a pixel-cell boundary (SURVEY §7 hard part 5)
    err = np.abs(Jn - Ja).reshape(len(Jn), -1).max(1)
    scale = np.abs(Ja).reshape(len(Ja), -1).max(1) + 1.0
    assert (err / scale < 1e-5).mean() > 0.98
    rp, Jp = oracle.eval_text(prob, TX_POSE, JAC_CENTRAL_DIFF)
    assert np.array_equal(rp, rn) and np.array_equal(Jp, Jn[:, :, :6])
    rt, Jt = oracle.eval_text(prob, TX_THETA, JAC_CENTRAL_DIFF)
    assert np.allclose(rt * prob.w_text, rn, rtol=1e-13)
    assert np.allclose(Jt * prob.w_text, Jn[:, :, 12:], rtol=1e-6, atol=1e-4)  # h ~ 1e-8: cancellation noise


def test_text_out_of_image_and_sigma_zero(oracle):
    prob = synth.c4_local_ba(seed=6, n_lm=20, n_planes=4)
    prob.t_musigma[:25, 1] = 0.0           # sigma == 0 -> residual 0 (nume_BAText.h:85-90)
    prob.t_rays[25:50] += 50.0             # projects far outside -> intensity 0 (:71-72)
    r, J = oracle.eval_text(prob, TX_BA, JAC_ANALYTIC)
    assert np.all(r[:25] == 0) and np.all(J[:25] == 0)
    mu, sg = prob.t_musigma[25, 0], prob.t_musigma[25, 1]
    assert np.allclose(r[25:50], ((0 - mu) / sg - prob.t_iref[25:50]) * prob.w_text)
    assert np.all(J[25:50] == 0)


def test_lm_converges_to_ground_truth(oracle):
    prob = synth.make_ba_problem(seed=1, n_kf=10, n_lm=600, obs_per_lm=3, band=10, fixed_cams=(0, 1, 2),
                                 pix_noise=0.0, outlier_frac=0.0)
    s, fr, tr = oracle.solve(prob, 30)
    cg, rg, _ = prob.gt
    assert s["final_cost"] < 1e-9 * s["initial_cost"]
    assert np.abs(prob.cams - cg).max() < 1e-6 and np.abs(prob.rho - rg).max() < 1e-6
    assert s["termination"] in (1, 2, 3)


def test_schur_equals_full_dense_solve(oracle):
    prob = synth.c4_local_ba(seed=7, n_lm=150, n_planes=6)
    a, b = prob.copy(), prob.copy()
    sa, fa, ta = oracle.solve(a, 6, JAC_CENTRAL_DIFF)
    sb, fb, tb = oracle.solve(b, 6, JAC_CENTRAL_DIFF, dense_full=1)
    assert sa["iterations"] == sb["iterations"]
    assert np.allclose(ta[: sa["iterations"] + 1, 0], tb[: sb["iterations"] + 1, 0], rtol=1e-8)
    assert np.allclose(a.cams, b.cams, atol=1e-7) and np.allclose(a.rho, b.rho, atol=1e-6)
    assert np.allclose(a.theta, b.theta, atol=1e-6)


def test_fixed_blocks_and_fixed_cost(oracle):
    prob = synth.make_ba_problem(seed=8, n_kf=6, n_lm=200, obs_per_lm=3, band=6, fixed_cams=(0, 1, 2), n_ext=3,
                                 frac_ext_lm=0.4, n_planes=4)
    before = prob.params()
    s, fr, tr = oracle.solve(prob, 5, JAC_ANALYTIC)
    fixed_c = prob.cam_fixed == 1
    assert np.array_equal(prob.cams[fixed_c], before[0][fixed_c])
    assert np.array_equal(prob.rho[prob.rho_fixed == 1], before[1][prob.rho_fixed == 1])
    assert np.array_equal(prob.theta[prob.theta_fixed == 1], before[2][prob.theta_fixed == 1])
    assert s["fixed_cost"] > 0  # blocks whose every parameter block is constant (Appendix A.7)
    assert s["final_cost"] < s["initial_cost"]


def test_huber_corrected_final_residuals(oracle):
    prob = synth.c3_pose_only(seed=9, n_pobs=300, n_planes=2)
    s, fr, tr = oracle.solve(prob, 3, JAC_ANALYTIC)
    r, _ = oracle.eval_points(prob, PT_POSE, want_J=False)
    sn = (r ** 2).sum(1)
    a = prob.huber_point
    scale = np.where(sn <= a * a, 1.0, np.sqrt(a / np.sqrt(np.maximum(sn, 1e-300))))
    assert np.allclose(fr[: 2 * prob.n_pobs].reshape(-1, 2), r * scale[:, None], rtol=1e-12, atol=1e-14)


def test_init_ba_functor_is_ba_nw_with_identity_host(oracle):
    """auto_IniBAScene (include/auto_IniBAScene.h:28-60) = rotate ray/rho by the second frame's pose, project, unweighted.
    With the host camera held at identity the auto_BASceneNW path must give the same residuals, so InitBA needs no extra kernel."""
    rng = np.random.default_rng(5)
    n = 200
    q = rng.normal(size=4); q /= np.linalg.norm(q) * 0.9     # deliberately not unit: QuaternionRotatePoint normalises
    q = np.array([1.0, 0.02, -0.03, 0.01]) * 1.1
    t = np.array([0.2, -0.1, 0.05])
    cams = np.stack([np.array([1.0, 0, 0, 0, 0, 0, 0]), np.concatenate([q, t])])
    ray = rng.uniform(-0.5, 0.5, (n, 2)); rho = rng.uniform(0.1, 1.0, n); uv = rng.uniform(0, 480, (n, 2))
    from textslam_b200 import BAProblem
    prob = BAProblem(cams, [1, 0], rho, None, None, None, uv, ray, np.ones(n, np.int32), np.zeros(n, np.int32), np.arange(n, dtype=np.int32),
                     w_point=(1.0, 1.0))
    r, _ = oracle.eval_points(prob, PT_BA_NW)
    fx, fy, cx, cy = prob.K_point
    un = q / np.linalg.norm(q)
    w, x, y, z = un
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    p = np.concatenate([ray, np.ones((n, 1))], 1) / rho[:, None]
    qp = p @ R.T + t
    want = np.stack([fx * qp[:, 0] / qp[:, 2] + cx - uv[:, 0], fy * qp[:, 1] / qp[:, 2] + cy - uv[:, 1]], 1)
    assert np.allclose(r, want, rtol=1e-12, atol=1e-9)
