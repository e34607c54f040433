"""Second, independent restatement of the residual functors in plain numpy, written from the reference headers
(include/auto_BAScene.h:27-87, include/nume_BAText.h:28-94, include/ModelTool.hpp:164-171, include/rotation.h:524-573) with
4x4 matrices / explicit quaternion algebra — no code shared with oracle/ba_math.hpp. The C++ oracle must give the same residuals to
rounding and, for its Jacobians (dual numbers / closed forms / Ceres-style central differences), the same derivatives as central
differences of THIS implementation taken through the quaternion Plus of ceres::QuaternionParameterization."""
import numpy as np
import pytest
from textslam_b200 import synth
from textslam_b200._abi import PT_BA, PT_BA_NW, PT_POSE, PT_RHO, TX_BA, JAC_ANALYTIC, JAC_CENTRAL_DIFF


def quat_to_R(q):
    w, x, y, z = q / np.sqrt((q * q).sum())     # QuaternionRotatePoint / Eigen::Quaterniond::normalized (no conjugate: complex-step safe)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def T44(cam):
    T = np.eye(4); T[:3, :3] = quat_to_R(cam[:4]); T[:3, 3] = cam[4:]
    return T


def point_residual(cam, host, rho, ray_xy, uv, K, w):
    """auto_BAScene: q_cr = q_cw (x) conj(q_rw); t_cr = t_cw - R(q_cr) t_rw; p = ray / rho; project. Written with matrices: the product of
    two unit rotations is the rotation of the quaternion product, but a NON-unit q_cr is normalised as a whole by QuaternionRotatePoint."""
    qc, qr = cam[:4], host[:4]
    qwr = np.array([qr[0], -qr[1], -qr[2], -qr[3]])
    a, b = qc, qwr
    qcr = np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                    a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])
    R = quat_to_R(qcr)
    tcr = cam[4:] - R @ host[4:]
    p = np.array([ray_xy[0], ray_xy[1], 1.0]) / rho
    X = R @ p + tcr
    fx, fy, cx, cy = K
    return np.array([(fx * X[0] / X[2] + cx - uv[0]) * w[0], (fy * X[1] / X[2] + cy - uv[1]) * w[1]])


def text_residual(cam, host, theta, rays, iref, mu, sigma, img, K, wT):
    """nume_BAText: Tcr = Tcw Trw^-1; TextProj; bilinear 2x2 of the u8 image with the reference's bounds test."""
    Tcr = T44(cam) @ np.linalg.inv(T44(host))
    fx, fy, cx, cy = K
    rows, stride = img.shape
    out = np.zeros(8)
    for i in range(8):
        ray = np.array([rays[i, 0], rays[i, 1], 1.0])
        rho = -ray @ theta
        p = Tcr[:3, :3] @ ray / rho + Tcr[:3, 3]
        u, v = fx * p[0] / p[2] + cx, fy * p[1] / p[2] + cy
        uf, vf, uc, vc = int(np.floor(u)), int(np.floor(v)), int(np.ceil(u)), int(np.ceil(v))
        if uf < 0 or vf < 0 or uc >= stride or vc >= rows:
            cur = 0.0
        else:
            su, sv = u - uf, v - vf
            cur = ((1 - su) * (1 - sv) * img[vf, uf] + su * (1 - sv) * img[vf, uf + 1] + (1 - su) * sv * img[vf + 1, uf] + su * sv * img[vf + 1, uf + 1])
        out[i] = ((cur - mu) / sigma - iref[i]) * wT if sigma != 0 else 0.0
    return out


def plus(cam, d):
    """ceres::QuaternionParameterization::Plus on the quaternion (delta on the LEFT: q' = exp(d) (x) q), identity on the translation."""
    n = np.sqrt((d[:3] * d[:3]).sum())      # complex-step safe (analytic continuation, no conjugate)
    if n != 0:
        s = np.sin(n) / n
        e = np.array([np.cos(n), s * d[0], s * d[1], s * d[2]])
    else:
        e = np.array([1.0, 0, 0, 0])
    q = cam[:4]
    qq = np.array([e[0] * q[0] - e[1] * q[1] - e[2] * q[2] - e[3] * q[3], e[0] * q[1] + e[1] * q[0] + e[2] * q[3] - e[3] * q[2],
                   e[0] * q[2] - e[1] * q[3] + e[2] * q[0] + e[3] * q[1], e[0] * q[3] + e[1] * q[2] - e[2] * q[1] + e[3] * q[0]])
    return np.concatenate([qq, cam[4:] + d[3:]])


def test_point_functors_match_numpy_restatement(oracle):
    prob = synth.c4_local_ba(seed=17, n_lm=120, n_planes=0)
    ro, Jo = oracle.eval_points(prob, PT_BA)
    rn, _ = oracle.eval_points(prob, PT_BA_NW)
    rng = np.random.default_rng(0)
    for i in rng.choice(prob.n_pobs, 40, replace=False):
        cam, host, rho = prob.cams[prob.p_cam[i]], prob.cams[prob.p_host[i]], prob.rho[prob.p_lm[i]]
        r = point_residual(cam, host, rho, prob.p_ray[i], prob.p_uv[i], prob.K_point, prob.w_point)
        assert np.allclose(r, ro[i], rtol=1e-11, atol=1e-10)
        assert np.allclose(point_residual(cam, host, rho, prob.p_ray[i], prob.p_uv[i], prob.K_point, (1.0, 1.0)), rn[i], rtol=1e-11, atol=1e-10)
        # tangent-space Jacobian [d_cam(6) d_host(6) d_rho] by central differences of the numpy functor
        eps, J = 1e-6, np.zeros((2, 13))
        for k in range(13):
            d = np.zeros(13); d[k] = eps
            f = lambda s: point_residual(plus(cam, s * d[:6]), plus(host, s * d[6:12]), rho + s * d[12], prob.p_ray[i], prob.p_uv[i], prob.K_point, prob.w_point)
            J[:, k] = (f(1.0) - f(-1.0)) / (2 * eps)
        assert np.allclose(J, Jo[i], rtol=2e-5, atol=2e-5 * np.abs(Jo[i]).max())
    # the constant-block variants are column subsets of the same functor
    rp, Jp = oracle.eval_points(prob, PT_POSE)
    rr, Jr = oracle.eval_points(prob, PT_RHO)
    assert np.allclose(rp, ro, rtol=1e-12, atol=1e-12) and np.allclose(Jp, Jo[:, :, :6], rtol=1e-12, atol=1e-12)
    assert np.allclose(rr, rn, rtol=1e-12, atol=1e-12)


def test_non_unit_quaternions_are_normalised_like_the_reference(oracle):
    prob = synth.c4_local_ba(seed=18, n_lm=40, n_planes=2)
    prob.cams[:, :4] *= np.linspace(0.7, 1.4, len(prob.cams))[:, None]     # Ceres never renormalises the ambient quaternion
    ro, _ = oracle.eval_points(prob, PT_BA, want_J=False)
    for i in range(0, prob.n_pobs, 7):
        r = point_residual(prob.cams[prob.p_cam[i]], prob.cams[prob.p_host[i]], prob.rho[prob.p_lm[i]], prob.p_ray[i], prob.p_uv[i], prob.K_point, prob.w_point)
        assert np.allclose(r, ro[i], rtol=1e-11, atol=1e-10)
    rt, _ = oracle.eval_text(prob, TX_BA, JAC_ANALYTIC, want_J=False)
    for j in range(0, prob.n_tobs, 5):
        r = text_residual(prob.cams[prob.t_cam[j]], prob.cams[prob.t_host[j]], prob.theta[prob.t_plane[j]], prob.t_rays[j], prob.t_iref[j],
                          prob.t_musigma[j, 0], prob.t_musigma[j, 1], prob.imgs[prob.t_img[j]], prob.K_text, prob.w_text)
        assert np.allclose(r, rt[j], rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("level", [0, 1])
def test_text_functor_matches_numpy_restatement(oracle, level):
    prob = synth.c4_local_ba(seed=19, level=level, n_lm=30, n_planes=6)
    ro, Ja = oracle.eval_text(prob, TX_BA, JAC_ANALYTIC)
    _, Jc = oracle.eval_text(prob, TX_BA, JAC_CENTRAL_DIFF)
    rng = np.random.default_rng(1)
    for j in rng.choice(prob.n_tobs, 12, replace=False):
        cam, host, th = prob.cams[prob.t_cam[j]], prob.cams[prob.t_host[j]], prob.theta[prob.t_plane[j]]
        args = (prob.t_rays[j], prob.t_iref[j], prob.t_musigma[j, 0], prob.t_musigma[j, 1], prob.imgs[prob.t_img[j]], prob.K_text, prob.w_text)
        assert np.allclose(text_residual(cam, host, th, *args), ro[j], rtol=1e-10, atol=1e-9)
        # the image is piecewise bilinear: a small central difference that stays inside one pixel cell reproduces the analytic slope
        eps, J = 1e-7, np.zeros((8, 15))
        for k in range(15):
            d = np.zeros(15); d[k] = eps
            f = lambda s: text_residual(plus(cam, s * d[:6]), plus(host, s * d[6:12]), th + s * d[12:], *args)
            J[:, k] = (f(1.0) - f(-1.0)) / (2 * eps)
        scale = np.abs(Ja[j]).max() + 1e-12
        ok = np.abs(J - Ja[j]) <= 5e-4 * scale
        assert ok.mean() >= 0.95, (j, np.abs(J - Ja[j]).max() / scale)     # rows whose sample sits on a pixel boundary may differ
        assert np.abs(Jc[j] - Ja[j]).max() <= 0.2 * scale                  # Ceres' step (1e-6 relative) straddles cell borders more often


def ceres_central_numeric_jacobian(fun, x):
    """ceres::NumericDiffCostFunction<CENTRAL> with default options: h_j = max(|x_j| 1e-6, sqrt(DBL_EPSILON))."""
    J = np.zeros((8, len(x)))
    for j in range(len(x)):
        h = max(abs(x[j]) * 1e-6, np.sqrt(np.finfo(float).eps))
        xp, xm = x.copy(), x.copy()
        xp[j] += h; xm[j] -= h
        J[:, j] = (fun(xp) - fun(xm)) / (2 * h)
    return J


def quaternion_plus_jacobian(q):
    """d Plus(q, delta) / d delta at delta = 0 for ceres::QuaternionParameterization (4 x 3, Appendix A.1)."""
    x0, x1, x2, x3 = q
    return np.array([[-x1, -x2, -x3], [x0, x3, -x2], [-x3, x0, x1], [x2, -x1, x0]])


def test_numeric_diff_mode_is_ceres_step_rule_plus_local_parameterization(oracle):
    """The oracle's TSLAM_JAC_CENTRAL_DIFF Jacobian = Ceres' central differences on the 17 AMBIENT parameters of nume_BAText (4+3+4+3+3),
    quaternion blocks then multiplied by the 4x3 Plus Jacobian: identical arithmetic, so agreement is at rounding level."""
    prob = synth.c4_local_ba(seed=23, n_lm=30, n_planes=5)
    _, Jc = oracle.eval_text(prob, TX_BA, JAC_CENTRAL_DIFF)
    for j in range(0, prob.n_tobs, 9):
        cam, host, th = prob.cams[prob.t_cam[j]].copy(), prob.cams[prob.t_host[j]].copy(), prob.theta[prob.t_plane[j]].copy()
        args = (prob.t_rays[j], prob.t_iref[j], prob.t_musigma[j, 0], prob.t_musigma[j, 1], prob.imgs[prob.t_img[j]], prob.K_text, prob.w_text)
        x = np.concatenate([cam, host, th])
        fun = lambda v: text_residual(v[:7], v[7:14], v[14:], *args)
        Ja = ceres_central_numeric_jacobian(fun, x)                       # 8 x 17 ambient
        J = np.concatenate([Ja[:, 0:4] @ quaternion_plus_jacobian(cam[:4]), Ja[:, 4:7],
                            Ja[:, 7:11] @ quaternion_plus_jacobian(host[:4]), Ja[:, 11:14], Ja[:, 14:17]], axis=1)
        assert np.allclose(J, Jc[j], rtol=1e-6, atol=1e-6 * (np.abs(Jc[j]).max() + 1e-12)), j
