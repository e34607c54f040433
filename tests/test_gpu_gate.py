"""GPU parity of the chi^2 gates (tslam_gate_residuals, tslam_solve_gated) and of the coarse-to-fine pyramid loop against the
CPU oracle: flags are integer outputs -> bit-exact; poses within the north-star tolerance (1e-5 relative)."""
import numpy as np
import pytest
import textslam_b200 as T
from textslam_b200 import synth
from textslam_b200.api import run_pyramid
from test_oracle_gate import random_gate_case, oracle_solve_gated, pose_levels

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n_p,n_obj", [(0, 500, 7), (1, 1, 1), (2, 100003, 300), (3, 0, 5), (4, 77, 0)])
def test_gate_residuals_bit_exact(ctx, oracle, seed, n_p, n_obj):
    fr, n_p, n_t, t_obj, sizes = random_gate_case(seed, n_p=n_p, n_obj=n_obj)
    for g in (T.gate_options(), T.gate_options(chi2_text=0.95, chi2_mono=5.991, w_point=(1.0, 0.7), w_text=1.0, text_ratio=0.3),
              T.gate_options(gate_points=False), T.gate_options(gate_text=False)):
        got = ctx.gate_residuals(fr, n_p, n_t, g, t_obj, sizes)
        want = oracle.gate_residuals(fr, n_p, n_t, g, t_obj, sizes)
        for a, b in zip(got[:3], want[:3]):
            assert np.array_equal(a, b)
        assert got[3] == want[3]


def test_gate_threshold_ties(ctx, oracle):
    """Residuals sitting exactly on / one ulp around the thresholds: the division and the product must round like the host's."""
    w = 1.0 / 1.2
    g = T.gate_options(relax_below_text_blocks=0)
    base = 3.5 * w
    vals = np.array([base, np.nextafter(base, 0), np.nextafter(base, 10), -base, np.nextafter(-base, 0), np.nextafter(-base, -10)])
    fr_p = np.stack([vals, np.zeros_like(vals)], 1).ravel()
    tb = 0.5 * 5.0
    tv = np.array([tb, np.nextafter(tb, 0), np.nextafter(tb, 10), -tb])
    rt = np.zeros((len(tv), 8)); rt[np.arange(len(tv)), np.arange(len(tv))] = tv
    fr = np.concatenate([fr_p, rt.ravel()])
    t_obj = np.zeros(len(tv), np.int32); sizes = np.array([len(tv)], np.int32)
    got = ctx.gate_residuals(fr, len(vals), len(tv), g, t_obj, sizes)
    want = oracle.gate_residuals(fr, len(vals), len(tv), g, t_obj, sizes)
    for a, b in zip(got[:3], want[:3]):
        assert np.array_equal(a, b)


def test_gate_rejects_inconsistent_bookkeeping(ctx):
    g = T.gate_options()
    with pytest.raises(T.TslamError):
        ctx.gate_residuals(np.zeros(24), 0, 3, g, np.array([0, 0, 0], np.int32), np.array([2], np.int32))
    with pytest.raises(T.TslamError):
        ctx.gate_residuals(np.zeros(8), 0, 1, g, np.array([3], np.int32), np.array([1], np.int32))
    # empty input is fine
    pb, tb, ob, cnt = ctx.gate_residuals(np.zeros(0), 0, 0, g)
    assert pb.size == 0 and tb.size == 0 and cnt == (0, 0, 0)


@pytest.mark.parametrize("maker,its", [(lambda: synth.c3_pose_only(seed=21, n_pobs=600, n_planes=5), 10),
                                       (lambda: synth.c4_local_ba(seed=22, n_lm=300, n_planes=8), 10)])
def test_solve_gated_equals_solve_then_oracle_gate(ctx, oracle, maker, its):
    prob = maker()
    n_obj = len(prob.theta)
    t_obj = prob.t_plane.astype(np.int32)
    sizes = np.bincount(t_obj, minlength=n_obj).astype(np.int32)
    g = T.gate_options(w_point=prob.w_point, w_text=prob.w_text)
    a, b = prob.copy(), prob.copy()
    summ, fr, _, pb, tb, ob, cnt = ctx.solve_gated(a, g, t_obj, sizes, its)
    s2, fr2, _ = ctx.solve(b, its)
    assert summ["iterations"] == s2["iterations"] and np.array_equal(fr, fr2) and np.array_equal(a.cams, b.cams)
    want = oracle.gate_residuals(fr2, prob.n_pobs, prob.n_tobs, g, t_obj, sizes)
    assert np.array_equal(pb, want[0]) and np.array_equal(tb, want[1]) and np.array_equal(ob, want[2]) and cnt == want[3]
    assert cnt[0] > 0   # the planted outliers are found
    # final_residuals may be omitted: flags must not change
    c = prob.copy()
    opt_none = ctx.solve_gated(c, g, t_obj, sizes, its, want_trace=False)
    assert np.array_equal(opt_none[3], pb) and np.array_equal(opt_none[4], tb)


def test_pose_pyramid_matches_oracle(ctx, oracle):
    """optimizer::PoseOptim's three-level loop (src/optimizer.cc:172-186) through the Optimizer mirror vs the same loop on the oracle."""
    n_obj = 4
    flags = lambda n: (np.ones(n, bool), np.ones(n_obj, bool), np.ones((n_obj, 25), bool))
    lg = pose_levels(seed=11); lo = pose_levels(seed=11)
    fg, fo = flags(lg[0].prob.n_pobs), flags(lo[0].prob.n_pobs)
    rg = T.Optimizer(ctx).PoseOptim(lg, 10, *fg)
    ro = run_pyramid(oracle_solve_gated(oracle), lo, (12.25,) * 3, (0.5, 0.5, 0.95), (10,) * 3, *fo)
    for a, b in zip(rg, ro):
        assert a["summary"]["iterations"] == b["summary"]["iterations"] and a["bad"] == b["bad"]
        assert a["n_point_blocks"] == b["n_point_blocks"] and a["n_text_blocks"] == b["n_text_blocks"]
    for x, y in zip(fg, fo):
        assert np.array_equal(x, y)
    assert np.abs(lg[0].prob.cams - lo[0].prob.cams).max() <= 1e-5 * np.abs(lo[0].prob.cams).max()


def test_local_ba_pyramid_matches_oracle(ctx, oracle):
    """optimizer::LocalBundleAdjustment's three PyrBA levels (src/optimizer.cc:282-289) through the Optimizer mirror vs the oracle:
    same blocks assembled per level, same flags cleared, poses / inverse depths within the north-star tolerance."""
    from textslam_b200.api import PyramidLevel
    n_obj = 6

    def levels():
        out = []
        for lv in (2, 1, 0):
            p = synth.c4_local_ba(seed=33, level=lv, n_lm=300, n_planes=n_obj)
            out.append(PyramidLevel(p, t_obj=p.t_plane.copy(), t_feat=np.tile(np.arange(25), n_obj)))
        return out

    lg, lo = levels(), levels()
    flags = lambda n: (np.ones(n, bool), np.ones(n_obj, bool), np.ones((n_obj, 25), bool))
    fg, fo = flags(lg[0].prob.n_pobs), flags(lo[0].prob.n_pobs)
    rg = T.Optimizer(ctx).LocalBundleAdjustment(lg, 10, *fg)
    ro = run_pyramid(oracle_solve_gated(oracle), lo, (12.25,) * 3, (0.5, 0.5, 0.95), (10,) * 3, *fo)
    for a, b in zip(rg, ro):
        assert a["summary"]["iterations"] == b["summary"]["iterations"] and a["bad"] == b["bad"]
        assert a["n_point_blocks"] == b["n_point_blocks"] and a["n_text_blocks"] == b["n_text_blocks"]
    for x, y in zip(fg, fo):
        assert np.array_equal(x, y)
    assert np.abs(lg[0].prob.cams - lo[0].prob.cams).max() <= 1e-5 * np.abs(lo[0].prob.cams).max()
    assert np.abs(lg[0].prob.rho - lo[0].prob.rho).max() <= 1e-5 * np.abs(lo[0].prob.rho).max()
