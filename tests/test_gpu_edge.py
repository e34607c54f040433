"""Edge cases of the solver boundary on the GPU: empty and fully constant problems, long landmark tracks, a larger camera
count than the benchmark shape (deeper nested-dissection tree, bigger block table)."""
import numpy as np
import pytest
from textslam_b200 import synth
from test_gpu_solve import _compare

pytestmark = pytest.mark.gpu


def _drop_observations(prob):
    for f in ("p_uv", "p_ray", "p_cam", "p_host", "p_lm", "t_rays", "t_iref", "t_musigma", "t_cam", "t_host", "t_plane", "t_img"):
        a = getattr(prob, f)
        setattr(prob, f, a[:0].copy())
    return prob


def test_problem_without_observations(ctx, oracle):
    prob = _drop_observations(synth.c4_local_ba(seed=3))
    before = prob.params()
    summ, fr, _ = ctx.solve(prob, 10)
    assert summ["iterations"] == 0 and summ["n_free_cams"] == 0 and summ["reduced_dim"] == 0
    assert fr.size == 0
    for x, y in zip(before, prob.params()):
        assert np.array_equal(x, y)
    assert ctx.compare_analysis(prob) == []


def test_everything_constant(ctx, oracle):
    prob = synth.make_ba_problem(seed=4, n_kf=5, n_lm=100, obs_per_lm=3, band=4, fixed_cams=(0, 1, 2, 3, 4), n_planes=3, feats_per_plane=4)
    prob.rho_fixed = np.ones(len(prob.rho), np.uint8)
    prob.theta_fixed = np.ones(len(prob.theta), np.uint8)
    before = prob.params()
    so, _, _ = oracle.solve(prob.copy(), 10)
    summ, fr, _ = ctx.solve(prob, 10)
    assert summ["iterations"] == so["iterations"] == 0 and summ["termination"] == so["termination"]
    assert np.isclose(summ["initial_cost"], so["initial_cost"], rtol=1e-10) and np.isclose(summ["final_cost"], so["final_cost"], rtol=1e-10)
    assert np.isclose(summ["fixed_cost"], so["fixed_cost"], rtol=1e-10) and summ["fixed_cost"] > 0   # every block is a constant cost
    for x, y in zip(before, prob.params()):
        assert np.array_equal(x, y)


def test_long_tracks(ctx, oracle):
    """A few landmarks seen from almost every keyframe (slot lists of ~40 cameras, ~800 Schur pairs per landmark)."""
    prob = synth.make_ba_problem(seed=6, n_kf=40, n_lm=60, obs_per_lm=38, band=40, fixed_cams=(0,), n_planes=2, feats_per_plane=4)
    assert ctx.compare_analysis(prob) == []
    _compare(ctx, oracle, prob, 6)


def test_thousand_keyframes(ctx, oracle):
    """Twice the cameras of C5: reduced system 5988 (94 tile rows), deeper nested-dissection tree, 1M-entry block table."""
    prob = synth.make_ba_problem(seed=8, n_kf=1000, n_lm=12000, obs_per_lm=4, band=10, fixed_cams=(0, 1), w_point=1.0, huber_point=np.sqrt(5.991))
    assert ctx.compare_analysis(prob) == []
    import textslam_b200 as T
    info = T.analyze_structure(prob)
    assert info["n_waves"] <= 20 < info["n_tiles"]
    _compare(ctx, oracle, prob, 4, n_threads=16)
