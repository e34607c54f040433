"""Generates tests/golden/ba_numpy_golden.npz: bundle-adjustment golden vectors produced by the INDEPENDENT numpy restatements
(tests/test_oracle_functors_py.py: functors from the reference headers; tests/test_oracle_lm_py.py: dense reading of the Ceres LM loop),
not by the C++ oracle and not by the CUDA path. Both of those are then checked against this file (tests/test_golden_ba.py)."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from textslam_b200 import synth                                                   # noqa: E402
from test_oracle_functors_py import point_residual, text_residual, plus           # noqa: E402
from test_oracle_lm_py import dense_ceres_lm, dense_ceres_lm_ba                   # noqa: E402

FIELDS = ("cams", "cam_fixed", "rho", "rho_fixed", "p_uv", "p_ray", "p_cam", "p_host", "p_lm")


def pack(prefix, prob, out):
    for f in FIELDS:
        out[prefix + f] = getattr(prob, f)
    out[prefix + "K"] = np.array(prob.K_point); out[prefix + "w"] = np.array(prob.w_point); out[prefix + "huber"] = np.array(prob.huber_point)


def main():
    out = {}
    # (1) residuals + tangent Jacobians (complex step through Plus) of auto_BAScene on a small local-BA shaped problem
    prob = synth.make_ba_problem(seed=101, n_kf=5, n_lm=40, obs_per_lm=3, band=5, fixed_cams=(0, 1))
    pack("e_", prob, out)
    n, h = prob.n_pobs, 1e-30
    r = np.zeros((n, 2)); J = np.zeros((n, 2, 13))
    for i in range(n):
        cam, host, rho = prob.cams[prob.p_cam[i]], prob.cams[prob.p_host[i]], prob.rho[prob.p_lm[i]]
        r[i] = point_residual(cam, host, rho, prob.p_ray[i], prob.p_uv[i], prob.K_point, prob.w_point)
        for k in range(13):
            d = np.zeros(13, dtype=complex); d[k] = 1j * h
            J[i, :, k] = point_residual(plus(cam.astype(complex), d[:6]), plus(host.astype(complex), d[6:12]), rho + d[12], prob.p_ray[i], prob.p_uv[i],
                                        prob.K_point, prob.w_point).imag / h
    out["e_r"], out["e_J"] = r, J
    # (2) pose-only LM trace (hard start: ends with rejected steps)
    for tag, seed, noise in (("p1_", 7, 0.6), ("p2_", 4, 8e-2)):
        prob = synth.make_ba_problem(seed=seed, n_kf=1, n_lm=150, obs_per_lm=1, band=20, fixed_cams=(), n_ext=20, frac_ext_lm=1.0, rot_noise=noise, trans_noise=noise)
        pack(tag, prob, out)
        blocks = {"K": prob.K_point, "w": prob.w_point,
                  "obs": [(int(prob.p_cam[i]), int(prob.p_host[i]), float(prob.rho[prob.p_lm[i]]), prob.p_ray[i], prob.p_uv[i]) for i in range(prob.n_pobs)]}
        tr, x = dense_ceres_lm(blocks, prob.cams, 0, prob.huber_point, 10)
        out[tag + "trace"], out[tag + "cam0"] = tr, x
    # (3) small BA LM trace (free cameras + inverse depths, dense, no Schur)
    prob = synth.make_ba_problem(seed=12, n_kf=4, n_lm=24, obs_per_lm=3, band=4, fixed_cams=(0, 1), rot_noise=3e-2, trans_noise=3e-2, rho_noise=0.1)
    pack("b_", prob, out)
    tr, cams, rho = dense_ceres_lm_ba(prob, 10)
    out["b_trace"], out["b_out_cams"], out["b_out_rho"] = tr, cams, rho
    # (4) nume_BAText residuals on quarter-resolution images (level 2: 160 x 120), four planes
    prob = synth.make_ba_problem(seed=55, n_kf=4, n_lm=8, obs_per_lm=2, band=4, fixed_cams=(0,), n_planes=4, feats_per_plane=9, level=2)
    pack("t_", prob, out)
    for f in ("theta", "theta_fixed", "t_rays", "t_iref", "t_musigma", "t_cam", "t_host", "t_plane", "t_img", "imgs"):
        out["t_" + f] = getattr(prob, f)
    out["t_Kt"], out["t_wt"] = np.array(prob.K_text), np.array(prob.w_text)
    out["t_r"] = np.array([text_residual(prob.cams[prob.t_cam[j]], prob.cams[prob.t_host[j]], prob.theta[prob.t_plane[j]], prob.t_rays[j], prob.t_iref[j],
                                         prob.t_musigma[j, 0], prob.t_musigma[j, 1], prob.imgs[prob.t_img[j]], prob.K_text, prob.w_text)
                           for j in range(prob.n_tobs)])
    np.savez_compressed(os.path.join(HERE, "ba_numpy_golden.npz"), **out)
    print("written", {k: v.shape for k, v in out.items() if k.endswith(("trace", "_r", "_J"))})


if __name__ == "__main__":
    main()
