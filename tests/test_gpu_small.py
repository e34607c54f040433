"""The small-problem path (csrc/ba_small.cu: one persistent cooperative kernel per tslam_solve for pose-only tracking and
local-BA windows, src/optimizer.cc:1218, 1600) against the general path (TSLAM_SMALL=0: structure analysis + one launch per
phase) and against the CPU oracle: same iteration counts, accept / reject sequence and termination, costs and parameters to the
north-star tolerance; exactly one kernel launch per solve."""
import os

import numpy as np
import pytest

import textslam_b200 as T
from textslam_b200 import synth
from textslam_b200._lib import lib

pytestmark = pytest.mark.gpu


def _both(ctx, prob, iters, **kw):
    out = []
    for flag in ("1", "0"):
        os.environ["TSLAM_SMALL"] = flag
        try:
            q = prob.copy()
            n0 = lib().tslam_launch_count()
            s, fr, tr = ctx.solve(q, iters, **kw)
            out.append((q, s, fr, tr, lib().tslam_launch_count() - n0))
        finally:
            os.environ.pop("TSLAM_SMALL", None)
    return out


def rel(x, y):
    return float(np.abs(x - y).max() / (np.abs(y).max() + 1e-300)) if y.size else 0.0


CASES = {
    "c3_pose_only": lambda: (synth.c3_pose_only(seed=101), 10),
    "c4_local_ba": lambda: (synth.c4_local_ba(seed=102), 10),
    "local_ba_points_only": lambda: (synth.c4_local_ba(seed=103, n_planes=0), 10),
    "local_ba_external_hosts": lambda: (synth.make_ba_problem(seed=104, n_kf=8, n_lm=400, obs_per_lm=3, band=8, fixed_cams=(0, 1, 2), n_ext=4,
                                                              frac_ext_lm=0.4, n_planes=8), 10),
    "ten_free_cameras": lambda: (synth.make_ba_problem(seed=105, n_kf=12, n_lm=600, obs_per_lm=4, band=12, fixed_cams=(0, 1), n_planes=10), 10),
    "one_observation": lambda: (synth.make_ba_problem(seed=106, n_kf=3, n_lm=1, obs_per_lm=1, band=3, fixed_cams=(0,), n_planes=0), 10),
    "converges_before_the_limit": lambda: (synth.c3_pose_only(seed=107, n_pobs=300, n_planes=2), 50),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_small_path_matches_general_path_and_oracle(ctx, oracle, name):
    prob, iters = CASES[name]()
    (qs, ss, frs, trs, ls), (qg, sg, frg, trg, lg) = _both(ctx, prob, iters)
    assert ls == 1 and lg > 5, (ls, lg)                      # one persistent kernel vs the per-phase launches
    for k in ("iterations", "successful_steps", "unsuccessful_steps", "termination", "n_free_cams", "reduced_dim"):
        assert ss[k] == sg[k], (k, ss, sg)
    n = ss["iterations"] + 1
    assert np.array_equal(trs[:n, 3], trg[:n, 3])
    assert np.allclose(trs[:n, 0], trg[:n, 0], rtol=1e-9, atol=1e-9 * trg[0, 0]) and np.allclose(trs[:n, 1], trg[:n, 1], rtol=1e-6)
    assert abs(ss["final_cost"] - sg["final_cost"]) <= 1e-9 * sg["final_cost"] + 1e-12 * sg["initial_cost"]
    assert abs(ss["initial_cost"] - sg["initial_cost"]) <= 1e-12 * sg["initial_cost"]
    assert rel(qs.cams, qg.cams) < 1e-8 and rel(qs.rho, qg.rho) < 1e-8 and rel(qs.theta, qg.theta) < 1e-8
    assert np.allclose(frs, frg, rtol=1e-7, atol=1e-7 * (np.abs(frg).max() + 1))
    qo = prob.copy()
    so, fo, to = oracle.solve(qo, iters)
    assert ss["iterations"] == so["iterations"] and ss["termination"] == so["termination"]
    assert abs(ss["final_cost"] - so["final_cost"]) <= 1e-8 * so["final_cost"] + 1e-12 * so["initial_cost"]
    assert rel(qs.cams, qo.cams) < 1e-5 and rel(qs.rho, qo.rho) < 1e-5 and rel(qs.theta, qo.theta) < 1e-5


def test_small_path_is_reproducible(ctx):
    """fixed-order sums everywhere: two runs give bit-identical results"""
    prob = synth.c4_local_ba(seed=108)
    a, b = prob.copy(), prob.copy()
    sa, fa, ta = ctx.solve(a, 10)
    sb, fb, tb = ctx.solve(b, 10)
    assert np.array_equal(a.cams, b.cams) and np.array_equal(a.rho, b.rho) and np.array_equal(a.theta, b.theta) and np.array_equal(fa, fb)
    assert sa["final_cost"] == sb["final_cost"]


def test_small_path_landmarks_only_and_theta(ctx, oracle):
    """every pose constant (OptimizeLandmarker / ThetaOptimMultiFs shapes, src/optimizer.cc:456-562, 2170-2242): no camera system at all"""
    prob = synth.c4_local_ba(seed=109, n_lm=200, n_planes=6)
    prob.cam_fixed[:] = 1
    (qs, ss, frs, trs, ls), (qg, sg, frg, trg, lg) = _both(ctx, prob, 50)
    assert ls == 1 and ss["n_free_cams"] == 0 and ss["iterations"] == sg["iterations"] and ss["termination"] == sg["termination"]
    assert abs(ss["final_cost"] - sg["final_cost"]) <= 1e-9 * sg["final_cost"]
    assert rel(qs.rho, qg.rho) < 1e-8 and rel(qs.theta, qg.theta) < 1e-8 and np.array_equal(qs.cams, prob.cams)


def test_small_path_huber_and_fixed_blocks(ctx, oracle):
    """robust loss on both block kinds, blocks whose parameters are all constant (they only add to the fixed cost), zero iterations"""
    prob = synth.c4_local_ba(seed=110, n_lm=300, n_planes=8)
    prob.huber_point, prob.huber_text = 2.0, 1.5
    prob.rho_fixed[:100] = 1
    prob.theta_fixed[:3] = 1
    (qs, ss, frs, trs, ls), (qg, sg, frg, trg, lg) = _both(ctx, prob, 10)
    assert ls == 1 and ss["iterations"] == sg["iterations"] and ss["termination"] == sg["termination"]
    assert abs(ss["final_cost"] - sg["final_cost"]) <= 1e-9 * sg["final_cost"] and abs(ss["fixed_cost"] - sg["fixed_cost"]) <= 1e-12 * (sg["fixed_cost"] + 1)
    assert rel(qs.cams, qg.cams) < 1e-8 and rel(qs.rho, qg.rho) < 1e-8 and rel(qs.theta, qg.theta) < 1e-8
    (q0, s0, fr0, tr0, l0), (q1, s1, fr1, tr1, l1) = _both(ctx, prob, 0)
    assert s0["iterations"] == 0 and s0["initial_cost"] == pytest.approx(s1["initial_cost"], rel=1e-12) and np.array_equal(q0.cams, prob.cams)
    assert np.allclose(fr0, fr1, rtol=1e-9, atol=1e-9)


def test_small_path_gated_pyramid(ctx, oracle):
    """the three-level PoseOptim loop with chi^2 gates (src/optimizer.cc:172-186): flags and poses equal on both paths"""
    from textslam_b200.api import PyramidLevel
    res = []
    for flag in ("1", "0"):
        os.environ["TSLAM_SMALL"] = flag
        try:
            prob = synth.c3_pose_only(seed=111, n_pobs=600, n_planes=5)
            rng = np.random.default_rng(5)
            prob.p_uv[rng.choice(prob.n_pobs, 40, replace=False)] += 25.0     # outliers for the gate
            key = np.stack([prob.t_cam, prob.t_plane], 1)
            first = np.r_[True, (key[1:] != key[:-1]).any(1)]
            t_obj = np.cumsum(first) - 1
            t_feat = np.arange(prob.n_tobs) - np.nonzero(first)[0][t_obj]
            pts = np.ones(prob.n_pobs, bool); tx = np.ones(int(t_obj.max()) + 1, bool); ft = np.ones((len(tx), int(t_feat.max()) + 1), bool)
            levels = [PyramidLevel(prob, np.arange(prob.n_pobs), t_obj, t_feat) for _ in range(3)]
            T.Optimizer(ctx).PoseOptim(levels, 10, pts, tx, ft)
            res.append((prob.cams.copy(), pts.copy(), tx.copy(), ft.copy()))
        finally:
            os.environ.pop("TSLAM_SMALL", None)
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])
    assert (~res[0][1]).sum() >= 30
    assert rel(res[0][0], res[1][0]) < 1e-8
